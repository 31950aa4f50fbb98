// Shared device/host helpers for the resdsic_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/resdsic_b200.h"

#define RDSIC_CHECK_ARG(cond) \
  do {                        \
    if (!(cond)) return RDSIC_E_ARG; \
  } while (0)

static inline int rdsic_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// SM count of the current device (idempotent per-device cache: racing host threads store the same value).
static inline int rdsic_sm_count() {
  static int cache[32] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool track = dev >= 0 && dev < 32;
  int sms = track ? cache[dev] : 0;
  if (!sms) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (track) cache[dev] = sms;
  }
  return sms;
}

// ---- elementwise math, written to track the reference's fp32 ATen ops -------
// nn.GELU() default: 0.5*x*(1+erf(x/sqrt(2)))
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_epilogue(int epi, float v, float res, float aux) {
  switch (epi) {
    case RDSIC_EPI_GELU: return gelu_erf(v);
    case RDSIC_EPI_RES_GELU: return gelu_erf(v + res);
    case RDSIC_EPI_ADD_RES: return v + res;
    case RDSIC_EPI_GATE: return aux * sigmoid_f(v) + res;
    case RDSIC_EPI_GDN: return res * rsqrtf(v);
    case RDSIC_EPI_IGDN: return res * sqrtf(v);
    case RDSIC_EPI_LRP: return res + 0.5f * tanhf(v);
    default: return v;
  }
}

__host__ __device__ static inline bool epi_needs_res(int epi) {
  return epi == RDSIC_EPI_RES_GELU || epi == RDSIC_EPI_ADD_RES || epi == RDSIC_EPI_GATE || epi == RDSIC_EPI_GDN ||
         epi == RDSIC_EPI_IGDN || epi == RDSIC_EPI_LRP;
}

// ---- typed element access through an rdsic_view -----------------------------
__device__ __forceinline__ float ld_elem(const void* p, int dtype, size_t i) {
  return dtype == RDSIC_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_elem(void* p, int dtype, size_t i, float v) {
  if (dtype == RDSIC_BF16)
    ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else
    ((float*)p)[i] = v;
}

// element index of (pixel, channel) in a view; `pix` = (b*H + y)*W + x, HW = H*W, C = channels (nchw only)
__device__ __forceinline__ size_t view_index(const rdsic_view& v, size_t pix, int c, int HW, int C) {
  if (v.nchw) {
    size_t b = pix / HW, r = pix % HW;
    return (b * C + c) * (size_t)HW + r;
  }
  return pix * (size_t)v.ld + v.coff + c;
}
