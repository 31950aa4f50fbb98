// Hyperprior entropy-model kernels (memory-bound, fp32 arithmetic in every mode).
//
//  eb_forward : EntropyBottleneck.forward eval (reference
//               entropy_models/entropy_models.py:447-490, _logits_cumulative
//               :401-420, _likelihood :422-434, quantize :139-148) fused with
//               the ste_round z_hat of models/WACNN/cnn.py:152-154.
//  gc_forward : GaussianConditional.forward eval + _likelihood (:627-661),
//               ste_round (ops/ops.py:34, cnn.py:177), quantize "symbols"
//               (:139-152) and build_indexes (:663-668) in ONE pass per slice
//               (the reference spends ~100 + 63 launches per slice on these).
//
// Floating-point op order follows the reference exactly:
// explicit __fmul_rn/__fadd_rn/__fsub_rn keep nvcc from contracting into FMAs
// where the reference rounds twice.
#include "common.cuh"

namespace {

// offsets inside one channel's RDSIC_EB_STRIDE-float parameter block
constexpr int EB_M0 = 0, EB_M1 = 3, EB_M2 = 12, EB_M3 = 21, EB_M4 = 30;
constexpr int EB_B0 = 33, EB_B1 = 36, EB_B2 = 39, EB_B3 = 42, EB_B4 = 45;
constexpr int EB_F0 = 46, EB_F1 = 49, EB_F2 = 52, EB_F3 = 55, EB_MED = 58;

__device__ __forceinline__ float eb_logits(const float* __restrict__ p, float v) {
  float a[3], t[3];
  // layer 0: [3x1] @ v + b, then += tanh(f)*tanh(.)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float l = __fadd_rn(__fmul_rn(p[EB_M0 + j], v), p[EB_B0 + j]);
    a[j] = __fadd_rn(l, __fmul_rn(p[EB_F0 + j], tanhf(l)));
  }
  const int mo[3] = {EB_M1, EB_M2, EB_M3}, bo[3] = {EB_B1, EB_B2, EB_B3}, fo[3] = {EB_F1, EB_F2, EB_F3};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* m = p + mo[k] + 3 * j;
      float l = __fmul_rn(m[0], a[0]);
      l = fmaf(m[1], a[1], l);
      l = fmaf(m[2], a[2], l);
      l = __fadd_rn(l, p[bo[k] + j]);
      t[j] = __fadd_rn(l, __fmul_rn(p[fo[k] + j], tanhf(l)));
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = t[j];
  }
  float l = __fmul_rn(p[EB_M4 + 0], a[0]);
  l = fmaf(p[EB_M4 + 1], a[1], l);
  l = fmaf(p[EB_M4 + 2], a[2], l);
  return __fadd_rn(l, p[EB_B4]);
}

__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__global__ void eb_forward_kernel(const rdsic_eb_desc d) {
  const size_t total = (size_t)d.B * d.h * d.w * d.C;
  const int hw = d.h * d.w;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % d.C);
    const size_t pix = e / d.C;
    const float* p = d.params + (size_t)c * RDSIC_EB_STRIDE;
    const float med = p[EB_MED];
    const float z = ((const float*)d.z.ptr)[pix * d.z.ld + d.z.coff + c];
    const float r = rintf(__fsub_rn(z, med));  // torch.round = half-to-even
    const float q = __fadd_rn(r, med);
    float ql = q;  // where the likelihood is evaluated: z_hat (eval) or z + noise (training)
    if (d.noise.ptr) {
      ql = __fadd_rn(z, ((const float*)d.noise.ptr)[pix * d.noise.ld + d.noise.coff + c]);
      if (d.noisy_out.ptr) ((float*)d.noisy_out.ptr)[pix * d.noisy_out.ld + d.noisy_out.coff + c] = ql;
    }
    const float lo = eb_logits(p, __fsub_rn(ql, 0.5f));
    const float up = eb_logits(p, __fadd_rn(ql, 0.5f));
    const float su = __fadd_rn(lo, up);
    const float sgn = su > 0.f ? -1.f : (su < 0.f ? 1.f : 0.f);  // -sign(lo+up)
    float lik = fabsf(__fsub_rn(sigmoid_ref(__fmul_rn(sgn, up)), sigmoid_ref(__fmul_rn(sgn, lo))));
    lik = fmaxf(lik, d.lik_bound);
    st_elem(d.z_hat.ptr, d.z_hat.dtype, pix * d.z_hat.ld + d.z_hat.coff + c, q);
    const size_t b = pix / hw, yx = pix % hw;
    const size_t o = (b * d.C + c) * (size_t)hw + yx;
    d.lik[o] = lik;
    if (d.symbols) d.symbols[o] = (int)r;
  }
}

// ------------------------------------------------------------------ GC
constexpr int GC_TP = 32;  // pixels per tile
constexpr int GC_TC = 32;  // channels per tile

__global__ void __launch_bounds__(256) gc_forward_kernel(const rdsic_gc_desc d) {
  __shared__ float s_lik[GC_TC][GC_TP + 1];
  __shared__ int s_sym[GC_TC][GC_TP + 1];
  __shared__ int s_idx[GC_TC][GC_TP + 1];
  __shared__ float s_tab[128];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  const int hw = d.h * d.w;
  const size_t npix = (size_t)d.B * hw;
  const size_t p0 = (size_t)blockIdx.x * GC_TP;
  const int c0 = blockIdx.y * GC_TC;
  const int nt = d.n_table - 1;  // entries compared (table[:-1])
  for (int i = threadIdx.x; i < nt; i += blockDim.x) s_tab[i] = d.table[i];
  __syncthreads();

  const float cst = -0.70710678118654752440f;  // float(-(2 ** -0.5))
  const int c = c0 + tx;
  for (int pp = ty; pp < GC_TP; pp += 8) {
    const size_t pix = p0 + pp;
    if (pix >= npix || c >= d.Cs) continue;
    const float y = ((const float*)d.y.ptr)[pix * d.y.ld + d.y.coff + c];
    const float mu = ((const float*)d.mu.ptr)[pix * d.mu.ld + d.mu.coff + c];
    const float sc = ((const float*)d.scale.ptr)[pix * d.scale.ld + d.scale.coff + c];
    const float r = rintf(__fsub_rn(y, mu));
    const float yh = __fadd_rn(r, mu);
    float yl = yh;  // where the likelihood is evaluated: y_hat (eval) or y + noise (training)
    if (d.noise.ptr) {
      yl = __fadd_rn(y, ((const float*)d.noise.ptr)[pix * d.noise.ld + d.noise.coff + c]);
      if (d.noisy_out.ptr) ((float*)d.noisy_out.ptr)[pix * d.noisy_out.ld + d.noisy_out.coff + c] = yl;
    }
    const float v = fabsf(__fsub_rn(yl, mu));
    const float s = fmaxf(sc, d.scale_bound);
    const float up = __fmul_rn(0.5f, erfcf(__fmul_rn(cst, __fdiv_rn(__fsub_rn(0.5f, v), s))));
    const float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(cst, __fdiv_rn(__fsub_rn(-0.5f, v), s))));
    const float lik = fmaxf(__fsub_rn(up, lo), d.lik_bound);
    // idx = #{t in table[:-1] : t < s}  ==  (n-1) - #{t : s <= t}; table ascending
    int lo_i = 0, hi_i = nt;
    while (lo_i < hi_i) {
      const int mid = (lo_i + hi_i) >> 1;
      if (s_tab[mid] < s) lo_i = mid + 1; else hi_i = mid;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (d.y_hat[k].ptr) st_elem(d.y_hat[k].ptr, d.y_hat[k].dtype, pix * d.y_hat[k].ld + d.y_hat[k].coff + c, yh);
    s_lik[tx][pp] = lik;
    s_sym[tx][pp] = (int)r;
    s_idx[tx][pp] = lo_i;
  }
  __syncthreads();
  // transposed write-out: tx = pixel (contiguous in NCHW), ty strides channels
  const size_t pix = p0 + tx;
  if (pix < npix) {
    const size_t b = pix / hw, yx = pix % hw;
    for (int cc = ty; cc < GC_TC; cc += 8) {
      const int ch = c0 + cc;
      if (ch >= d.Cs) break;
      const size_t o = (b * d.Ctot + d.lik_coff + ch) * (size_t)hw + yx;
      d.lik[o] = s_lik[cc][tx];
      if (d.symbols) d.symbols[o] = s_sym[cc][tx];
      if (d.indexes) d.indexes[o] = s_idx[cc][tx];
    }
  }
}

// EntropyBottleneck.loss: one thread per (channel, quantile); fixed-order block tree sum.
__global__ void __launch_bounds__(1024) eb_aux_loss_kernel(const float* __restrict__ params, const float* __restrict__ quantiles,
                                                          const float* __restrict__ target, int C, float* __restrict__ terms,
                                                          float* __restrict__ sum) {
  __shared__ float s_part[1024];
  float acc = 0.f;
  for (int e = threadIdx.x; e < C * 3; e += blockDim.x) {  // strided, so the order is fixed for a given C
    const int c = e / 3, k = e % 3;
    const float t = fabsf(__fsub_rn(eb_logits(params + (size_t)c * RDSIC_EB_STRIDE, quantiles[e]), target[k]));
    if (terms) terms[e] = t;
    acc = __fadd_rn(acc, t);
  }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s_part[threadIdx.x] = __fadd_rn(s_part[threadIdx.x], s_part[threadIdx.x + w]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *sum = s_part[0];
}

}  // namespace

extern "C" int rdsic_eb_aux_loss(const float* params, const float* quantiles, const float* target, int32_t C, float* terms,
                                 float* sum, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(params && quantiles && target && sum && C > 0);
  eb_aux_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(params, quantiles, target, C, terms, sum);
  return rdsic_launch_status();
}

extern "C" int rdsic_eb_forward(const rdsic_eb_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->z.ptr && d->z_hat.ptr && d->lik && d->params);
  RDSIC_CHECK_ARG(d->B > 0 && d->h > 0 && d->w > 0 && d->C > 0);
  RDSIC_CHECK_ARG(d->z.dtype == RDSIC_F32 && !d->z.nchw && !d->z_hat.nchw);
  RDSIC_CHECK_ARG(!d->noise.ptr || (d->noise.dtype == RDSIC_F32 && !d->noise.nchw));
  RDSIC_CHECK_ARG(!d->noisy_out.ptr || (d->noise.ptr && d->noisy_out.dtype == RDSIC_F32 && !d->noisy_out.nchw));
  const size_t total = (size_t)d->B * d->h * d->w * d->C;
  const int nblk = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  eb_forward_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(*d);
  return rdsic_launch_status();
}

extern "C" int rdsic_gc_forward(const rdsic_gc_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->y.ptr && d->mu.ptr && d->scale.ptr && d->lik && d->table);
  RDSIC_CHECK_ARG(d->B > 0 && d->h > 0 && d->w > 0 && d->Cs > 0 && d->Ctot >= d->lik_coff + d->Cs);
  RDSIC_CHECK_ARG(d->n_table >= 2 && d->n_table <= 129);
  RDSIC_CHECK_ARG(d->y.dtype == RDSIC_F32 && d->mu.dtype == RDSIC_F32 && d->scale.dtype == RDSIC_F32);
  RDSIC_CHECK_ARG(!d->y.nchw && !d->mu.nchw && !d->scale.nchw);
  RDSIC_CHECK_ARG(!d->noise.ptr || (d->noise.dtype == RDSIC_F32 && !d->noise.nchw));
  RDSIC_CHECK_ARG(!d->noisy_out.ptr || (d->noise.ptr && d->noisy_out.dtype == RDSIC_F32 && !d->noisy_out.nchw));
  const size_t npix = (size_t)d->B * d->h * d->w;
  dim3 grid((unsigned)((npix + GC_TP - 1) / GC_TP), (unsigned)ceil_div(d->Cs, GC_TC));
  gc_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d);
  return rdsic_launch_status();
}
