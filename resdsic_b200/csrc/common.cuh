// Shared device/host helpers for the resdsic_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A forward pass is ~300 dependent launches, a third of them 15-40 us long, so launch latency and the prologue (barrier
// init, TMEM allocation, tensor-map prefetch) of kernel n+1 looked worth hiding under kernel n's tail: every kernel of
// the forward path calls pdl_trigger() first thing -- the next kernel in the stream, if launched with the
// programmatic-serialization attribute (rdsic_launch below), may then become resident wherever an SM frees up and run
// its prologue -- and pdl_wait() before its first access to global memory, which blocks until every prerequisite grid
// has completed and flushed.  Both are no-ops for a launch without the attribute / without a dependent.
// MEASURED (batch 24, graph replay, A/B on one box): 1467-1476 images/s with the attribute vs 1481-1486 without, i.e.
// 0.8 % SLOWER -- the persistent one-CTA-per-SM kernels leave no room for a dependent CTA until their own CTA on that SM
// exits, so only the launch latency is hidden, and programmatic edges cost more than they save across the graph's
// parallel lanes.  Hence OFF by default; RDSIC_PDL=1 switches the attribute on.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static inline bool rdsic_pdl_enabled() {
  static const int on = getenv("RDSIC_PDL") ? atoi(getenv("RDSIC_PDL")) : 0;
  return on != 0;
}

// One launch path for the forward kernels: optional 2-CTA cluster, PDL attribute.
template <typename... KArgs, typename... Args>
static inline int rdsic_launch(void (*kern)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t stream, bool cluster2,
                               Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster2) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (rdsic_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return (int)e;
  return rdsic_launch_status();
}

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
