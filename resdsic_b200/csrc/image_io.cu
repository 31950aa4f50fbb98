// Host-facing image I/O of the codec's forward path (SURVEY 8d "end to end"; the reference's evaluation loop
// eval_model/__main__.py:133-147 moves fp32 images and full likelihood tensors across PCIe and reduces them on the host):
//
//   rdsic_image_u8_to_f32   8-bit image bytes -> x in [0,1] (torchvision ToTensor's arithmetic: v / 255)
//   rdsic_image_f32_to_u8   x_hat -> 8-bit: round(clamp(x_hat, 0, 1) * 255) (what an image writer stores)
//   rdsic_rate_per_image    bits[b] = -(sum log2 lik_y[b] + sum log2 lik_z[b])  (RateDistortionLoss's rate term per
//                           image, training/loss.py:14-22, before the division by the pixel count), fp64, fixed
//                           summation order (per-block partial sums, then one block per image adds them in index order)
//
// With them the device returns 1 byte per sample and 8 bytes per image instead of 4 bytes per sample plus both
// likelihood tensors: 28 MB instead of 162 MB per 24-image step at 512 x 768, and 28 MB instead of 113 MB going in.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n) {
  pdl_trigger();
  pdl_wait();
  const size_t n16 = n / 16, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(dst) + 4 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = make_float4((float)(w[k] & 0xFFu) / 255.0f, (float)((w[k] >> 8) & 0xFFu) / 255.0f,
                         (float)((w[k] >> 16) & 0xFFu) / 255.0f, (float)(w[k] >> 24) / 255.0f);
  }
  for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i] / 255.0f;
}

__device__ __forceinline__ uint32_t to_u8(float x) { return (uint32_t)__float2int_rn(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f); }

__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, size_t n) {
  pdl_trigger();
  pdl_wait();
  const size_t n16 = n / 16, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const float4* p = reinterpret_cast<const float4*>(src) + 4 * i;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 f = p[k];
      w[k] = to_u8(f.x) | (to_u8(f.y) << 8) | (to_u8(f.z) << 16) | (to_u8(f.w) << 24);
    }
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (uint8_t)to_u8(src[i]);
}

constexpr int RATE_BLOCKS = 64;  // partial sums per image and tensor

// grid (RATE_BLOCKS, B, 2): block (k, b, t) sums log2 of its contiguous slice of tensor t of image b
__global__ void __launch_bounds__(256) rate_partial_kernel(const float* __restrict__ lik_y, size_t ny, const float* __restrict__ lik_z,
                                                           size_t nz, double* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, t = blockIdx.z;
  const size_t n = t ? nz : ny;
  const float* p = (t ? lik_z : lik_y) + (size_t)b * n;
  const size_t per = (n + RATE_BLOCKS - 1) / RATE_BLOCKS, lo = (size_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  double acc = 0.0;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += (double)log2f(p[i]);
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[((size_t)b * 2 + t) * RATE_BLOCKS + blockIdx.x] = sh[0];
}

__global__ void rate_final_kernel(const double* __restrict__ partial, double* __restrict__ bits, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double acc = 0.0;
  for (int k = 0; k < 2 * RATE_BLOCKS; ++k) acc += partial[(size_t)b * 2 * RATE_BLOCKS + k];
  bits[b] = -acc;
}

}  // namespace

extern "C" int rdsic_image_u8_to_f32(const uint8_t* src, float* dst, size_t n, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(src && dst && n > 0);
  if (((uintptr_t)src % 16) || ((uintptr_t)dst % 16)) return RDSIC_E_ALIGN;
  const size_t want = (n / 16 + 255) / 256 + 1, cap = (size_t)rdsic_sm_count() * 8;
  return rdsic_launch(u8_to_f32_kernel, dim3((unsigned)(want < cap ? want : cap)), 256, 0, (cudaStream_t)stream, false, src, dst, n);
}

extern "C" int rdsic_image_f32_to_u8(const float* src, uint8_t* dst, size_t n, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(src && dst && n > 0);
  if (((uintptr_t)src % 16) || ((uintptr_t)dst % 16)) return RDSIC_E_ALIGN;
  const size_t want = (n / 16 + 255) / 256 + 1, cap = (size_t)rdsic_sm_count() * 8;
  return rdsic_launch(f32_to_u8_kernel, dim3((unsigned)(want < cap ? want : cap)), 256, 0, (cudaStream_t)stream, false, src, dst, n);
}

extern "C" int rdsic_rate_per_image(const float* lik_y, size_t ny, const float* lik_z, size_t nz, int B, double* workspace,
                                    double* bits, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(lik_y && lik_z && workspace && bits && B > 0 && B <= 65535 && ny > 0 && nz > 0);
  rate_partial_kernel<<<dim3(RATE_BLOCKS, (unsigned)B, 2), 256, 0, (cudaStream_t)stream>>>(lik_y, ny, lik_z, nz, workspace);
  rate_final_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(workspace, bits, B);
  return rdsic_launch_status();
}

extern "C" int rdsic_rate_workspace_doubles(int B) { return B > 0 ? B * 2 * RATE_BLOCKS : 0; }
