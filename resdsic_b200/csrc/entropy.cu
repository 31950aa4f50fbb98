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
#include <stdlib.h>

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

// One persistent block per SM.  Consecutive threads take consecutive channels (z is channels-last), so reading the
// 60 per-channel parameters straight from global memory makes every one of the ~120 loads per element touch 32
// different cache lines (the first version: 41 us for 0.44 M elements).  The table (C x 60 floats, 46 KB at C = 192)
// is therefore staged once per block in shared memory, at a row pitch of 61 words so that the 32 channels of a warp
// fall into 32 different banks.
constexpr int EB_PITCH = RDSIC_EB_STRIDE + 1;
__global__ void __launch_bounds__(256) eb_forward_kernel(const rdsic_eb_desc d) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_par[];  // [C][EB_PITCH]
  for (int i = threadIdx.x; i < d.C * RDSIC_EB_STRIDE; i += blockDim.x)
    s_par[(i / RDSIC_EB_STRIDE) * EB_PITCH + i % RDSIC_EB_STRIDE] = d.params[i];
  __syncthreads();
  const size_t total = (size_t)d.B * d.h * d.w * d.C;
  const int hw = d.h * d.w;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % d.C);
    const size_t pix = e / d.C;
    const float* p = s_par + (size_t)c * EB_PITCH;
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

// one element of GaussianConditional.forward + quantise + index build (the reference's fp32 op order)
struct GcOut {
  float yh, yl, lik;
  int sym, idx;
};
__device__ __forceinline__ GcOut gc_element(const rdsic_gc_desc& d, const float* s_tab, int nt, float y, float mu, float sc,
                                            float msk, float noise, bool has_mask, bool has_noise, bool from_sym, float r_in) {
  GcOut o;
  if (has_mask) sc = __fmul_rn(sc, msk);  // ResDSIC progressive stream: scale * mask, rint(y - mu) * mask + mu
  if (d.scale_eps != 0.f) sc = __fadd_rn(sc, d.scale_eps);  // cimd: scale * mask + 1e-7
  const float r = from_sym ? r_in : rintf(__fsub_rn(y, mu));
  const float rm = has_mask ? __fmul_rn(r, msk) : r;
  o.yh = __fadd_rn(rm, mu);
  o.yl = __fadd_rn(r, mu);  // where the likelihood is evaluated: round(y - mu) + mu (eval) or y + noise (training)
  if (has_noise) o.yl = __fadd_rn(y, noise);
  const float cst = -0.70710678118654752440f;  // float(-(2 ** -0.5))
  const float v = fabsf(__fsub_rn(o.yl, mu));
  const float s = fmaxf(sc, d.scale_bound);
  const float up = __fmul_rn(0.5f, erfcf(__fmul_rn(cst, __fdiv_rn(__fsub_rn(0.5f, v), s))));
  const float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(cst, __fdiv_rn(__fsub_rn(-0.5f, v), s))));
  o.lik = fmaxf(__fsub_rn(up, lo), d.lik_bound);
  // idx = #{t in table[:-1] : t < s}  ==  (n-1) - #{t : s <= t}; table ascending (only when the caller wants indexes:
  // the plain forward does not, and the search is ~15 % of the kernel's instructions)
  int lo_i = 0, hi_i = d.indexes ? nt : 0;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i) >> 1;
    if (s_tab[mid] < s) lo_i = mid + 1; else hi_i = mid;
  }
  o.sym = (int)rm;
  o.idx = lo_i;
  return o;
}

__global__ void __launch_bounds__(256) gc_forward_kernel(const rdsic_gc_desc d) {
  pdl_trigger();
  pdl_wait();
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

  const int c = c0 + tx;
  for (int pp = ty; pp < GC_TP; pp += 8) {
    const size_t pix = p0 + pp;
    if (pix >= npix || c >= d.Cs) continue;
    const float mu = ((const float*)d.mu.ptr)[pix * d.mu.ld + d.mu.coff + c];
    const float sc = ((const float*)d.scale.ptr)[pix * d.scale.ld + d.scale.coff + c];
    const float msk = d.mask.ptr ? ((const float*)d.mask.ptr)[pix * d.mask.ld + d.mask.coff + c] : 1.f;
    float y = 0.f, r = 0.f;
    if (d.sym_in) {  // decoder side: the symbol comes from the entropy decoder
      const size_t bb = pix / hw, yx = pix % hw;
      r = (float)d.sym_in[(bb * d.Ctot + d.lik_coff + c) * (size_t)hw + yx];
    } else {
      y = ((const float*)d.y.ptr)[pix * d.y.ld + d.y.coff + c];
    }
    const float nz = d.noise.ptr ? ((const float*)d.noise.ptr)[pix * d.noise.ld + d.noise.coff + c] : 0.f;
    const GcOut o = gc_element(d, s_tab, nt, y, mu, sc, msk, nz, d.mask.ptr != nullptr, d.noise.ptr != nullptr, d.sym_in != nullptr, r);
    if (d.noise.ptr && d.noisy_out.ptr) ((float*)d.noisy_out.ptr)[pix * d.noisy_out.ld + d.noisy_out.coff + c] = o.yl;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (d.y_hat[k].ptr) st_elem(d.y_hat[k].ptr, d.y_hat[k].dtype, pix * d.y_hat[k].ld + d.y_hat[k].coff + c, o.yh);
    s_lik[tx][pp] = o.lik;
    s_sym[tx][pp] = o.sym;
    s_idx[tx][pp] = o.idx;
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

// 128-bit form of the same kernel (the host selects it when every view allows it: channel counts / offsets / row
// pitches that are multiples of 4 floats, 16-byte aligned bases, h*w a multiple of 4).  A tile is still 32 pixels x
// 32 channels; on the way in a thread owns 4 consecutive CHANNELS of one pixel (one LDG.128 per operand, a warp
// reads four 128-byte pixel rows), on the way out -- through a shared-memory transpose -- 4 consecutive PIXELS of one
// channel (one STG.128 each for likelihood / symbols / indexes in the module's NCHW output).
constexpr int GC_VP = GC_TP + 4;  // transposed tile pitch: rows stay 16-byte aligned
__device__ __forceinline__ float4 ldg4(const rdsic_view& v, size_t pix, int c) {
  return __ldg(reinterpret_cast<const float4*>((const float*)v.ptr + pix * v.ld + v.coff + c));
}
__global__ void __launch_bounds__(256) gc_forward_vec_kernel(const rdsic_gc_desc d) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float s_lik[GC_TC][GC_VP];
  __shared__ __align__(16) int s_sym[GC_TC][GC_VP];
  __shared__ __align__(16) int s_idx[GC_TC][GC_VP];
  __shared__ float s_tab[128];
  const int hw = d.h * d.w;
  const size_t npix = (size_t)d.B * hw;
  const size_t p0 = (size_t)blockIdx.x * GC_TP;
  const int c0 = blockIdx.y * GC_TC;
  const int nt = d.n_table - 1;
  for (int i = threadIdx.x; i < nt; i += blockDim.x) s_tab[i] = d.table[i];
  __syncthreads();

  {
    const int cg = threadIdx.x % 8, pp = threadIdx.x / 8;  // 8 channel quads x 32 pixels
    const size_t pix = p0 + pp;
    const int c = c0 + 4 * cg;
    if (pix < npix && c < d.Cs) {
      const bool has_mask = d.mask.ptr != nullptr, has_noise = d.noise.ptr != nullptr, from_sym = d.sym_in != nullptr;
      const float4 mu4 = ldg4(d.mu, pix, c), sc4 = ldg4(d.scale, pix, c);
      float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f), m4 = make_float4(1.f, 1.f, 1.f, 1.f), n4 = y4, r4 = y4;
      if (from_sym) {
        const size_t bb = pix / hw, yx = pix % hw;
        const int32_t* sp = d.sym_in + (bb * d.Ctot + d.lik_coff + c) * (size_t)hw + yx;
        r4 = make_float4((float)sp[0], (float)sp[hw], (float)sp[2 * (size_t)hw], (float)sp[3 * (size_t)hw]);
      } else {
        y4 = ldg4(d.y, pix, c);
      }
      if (has_mask) m4 = ldg4(d.mask, pix, c);
      if (has_noise) n4 = ldg4(d.noise, pix, c);
      const float ys[4] = {y4.x, y4.y, y4.z, y4.w}, mus[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, scs[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
      const float ms[4] = {m4.x, m4.y, m4.z, m4.w}, ns[4] = {n4.x, n4.y, n4.z, n4.w}, rs[4] = {r4.x, r4.y, r4.z, r4.w};
      float yh[4], yl[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const GcOut o = gc_element(d, s_tab, nt, ys[i], mus[i], scs[i], ms[i], ns[i], has_mask, has_noise, from_sym, rs[i]);
        yh[i] = o.yh;
        yl[i] = o.yl;
        s_lik[4 * cg + i][pp] = o.lik;
        s_sym[4 * cg + i][pp] = o.sym;
        s_idx[4 * cg + i][pp] = o.idx;
      }
      if (has_noise && d.noisy_out.ptr)
        *reinterpret_cast<float4*>((float*)d.noisy_out.ptr + pix * d.noisy_out.ld + d.noisy_out.coff + c) = make_float4(yl[0], yl[1], yl[2], yl[3]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const rdsic_view& v = d.y_hat[k];
        if (!v.ptr) continue;
        const size_t e = pix * v.ld + v.coff + c;
        if (v.dtype == RDSIC_BF16) {
          __nv_bfloat162 a = __floats2bfloat162_rn(yh[0], yh[1]), b = __floats2bfloat162_rn(yh[2], yh[3]);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&a);
          u.y = *reinterpret_cast<uint32_t*>(&b);
          *reinterpret_cast<uint2*>((__nv_bfloat16*)v.ptr + e) = u;
        } else {
          *reinterpret_cast<float4*>((float*)v.ptr + e) = make_float4(yh[0], yh[1], yh[2], yh[3]);
        }
      }
    }
  }
  __syncthreads();
  {
    const int pg = threadIdx.x % 8, cc = threadIdx.x / 8;  // 8 pixel quads x 32 channels
    const size_t pix = p0 + 4 * pg;
    const int ch = c0 + cc;
    if (pix < npix && ch < d.Cs) {  // npix and p0 are multiples of 4: a quad is inside the tensor and inside one image
      const size_t b = pix / hw, yx = pix % hw;
      const size_t o = (b * d.Ctot + d.lik_coff + ch) * (size_t)hw + yx;
      *reinterpret_cast<float4*>(d.lik + o) = *reinterpret_cast<const float4*>(&s_lik[cc][4 * pg]);
      if (d.symbols) *reinterpret_cast<int4*>(d.symbols + o) = *reinterpret_cast<const int4*>(&s_sym[cc][4 * pg]);
      if (d.indexes) *reinterpret_cast<int4*>(d.indexes + o) = *reinterpret_cast<const int4*>(&s_idx[cc][4 * pg]);
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

// ------------------------------------------------------------------ CDF tables (update())
__global__ void gc_cdf_sizes_kernel(const float* __restrict__ table, int rows, float mult, int* __restrict__ offset,
                                    int* __restrict__ cdf_length) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const int center = (int)ceilf(__fmul_rn(table[i], mult));  // torch.ceil(scale_table * multiplier).int()
  offset[i] = -center;
  cdf_length[i] = 2 * center + 1 + 2;
}

// one block per table row; prob[row][j] = pmf(j), prob[row][pmf_length] = tail mass = 2 * lower(j = 0)
__global__ void gc_pmf_kernel(const float* __restrict__ table, const int* __restrict__ offset, float* __restrict__ prob, int ld) {
  const int row = blockIdx.x, center = -offset[row], len = 2 * center + 1;
  const float scale = table[row];
  const float cst = -0.70710678118654752440f;
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    const int d = j - center;
    const float s = (float)(d < 0 ? -d : d);
    const float up = __fmul_rn(0.5f, erfcf(__fmul_rn(cst, __fdiv_rn(__fsub_rn(0.5f, s), scale))));
    const float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(cst, __fdiv_rn(__fsub_rn(-0.5f, s), scale))));
    prob[(size_t)row * ld + j] = __fsub_rn(up, lo);
    if (j == 0) prob[(size_t)row * ld + len] = __fmul_rn(2.0f, lo);
  }
}

__global__ void eb_cdf_sizes_kernel(const float* __restrict__ quantiles, int C, int* __restrict__ offset, int* __restrict__ cdf_length) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float q0 = quantiles[3 * c], med = quantiles[3 * c + 1], q2 = quantiles[3 * c + 2];
  int minima = (int)ceilf(__fsub_rn(med, q0)), maxima = (int)ceilf(__fsub_rn(q2, med));
  minima = minima < 0 ? 0 : minima;
  maxima = maxima < 0 ? 0 : maxima;
  offset[c] = -minima;
  cdf_length[c] = maxima + minima + 1 + 2;
}

__device__ __forceinline__ float eb_pmf_at(const float* p, float v, float* lower, float* upper) {
  const float lo = eb_logits(p, __fsub_rn(v, 0.5f)), up = eb_logits(p, __fadd_rn(v, 0.5f));
  const float su = __fadd_rn(lo, up);
  const float sgn = su > 0.f ? -1.f : (su < 0.f ? 1.f : 0.f);
  *lower = lo;
  *upper = up;
  return fabsf(__fsub_rn(sigmoid_ref(__fmul_rn(sgn, up)), sigmoid_ref(__fmul_rn(sgn, lo))));
}

// one block per channel; samples j + (median - minima), j < max_length (entropy_models.py:373-388)
__global__ void eb_pmf_kernel(const float* __restrict__ params, const float* __restrict__ quantiles, const int* __restrict__ offset,
                              const int* __restrict__ cdf_length, int max_length, float* __restrict__ prob, int ld) {
  const int c = blockIdx.x, len = cdf_length[c] - 2;
  const float* p = params + (size_t)c * RDSIC_EB_STRIDE;
  const float start = __fsub_rn(quantiles[3 * c + 1], (float)(-offset[c]));
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    float lo, up;
    prob[(size_t)c * ld + j] = eb_pmf_at(p, __fadd_rn((float)j, start), &lo, &up);
  }
  if (threadIdx.x == 0) {  // tail mass = sigmoid(lower[0]) + sigmoid(-upper[max_length - 1])
    float lo0, up0, lo1, up1;
    eb_pmf_at(p, __fadd_rn(0.0f, start), &lo0, &up0);
    eb_pmf_at(p, __fadd_rn((float)(max_length - 1), start), &lo1, &up1);
    prob[(size_t)c * ld + len] = __fadd_rn(sigmoid_ref(lo0), sigmoid_ref(-up1));
  }
}

// pmf_to_quantized_cdf: one warp per row, the row's cdf in shared memory (see oracle/cdf_oracle.py for the
// algorithm and its provenance).  Integer arithmetic throughout after the fp32 round(p * 2^precision).
__global__ void __launch_bounds__(32) pmf_to_cdf_kernel(const float* __restrict__ prob, int ld, const int* __restrict__ cdf_length,
                                                        int precision, int* __restrict__ cdf_out, int cdf_ld, int* status) {
  extern __shared__ unsigned int s_cdf[];
  const int row = blockIdx.x, lane = threadIdx.x;
  const int n = cdf_length[row];  // cdf entries; n - 1 probabilities
  const float* p = prob + (size_t)row * ld;
  const unsigned int one = 1u << precision;
  bool bad = false;
  unsigned long long part = 0;
  if (lane == 0) s_cdf[0] = 0;
  for (int k = lane; k < n - 1; k += 32) {
    const float v = p[k];
    if (!(v >= 0.f) || isinf(v)) bad = true;
    const unsigned int f = bad ? 0u : (unsigned int)roundf(__fmul_rn(v, (float)one));  // std::round: half away from zero
    s_cdf[k + 1] = f;
    part += f;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  bad = __any_sync(0xffffffffu, bad) || part == 0 || n < 2;
  if (bad) {
    if (lane == 0) atomicCAS(status, 0, 1 + row);
    for (int k = lane; k < cdf_ld; k += 32) cdf_out[(size_t)row * cdf_ld + k] = 0;
    return;
  }
  const unsigned long long total = part;
  __syncwarp();
  // rescale, then inclusive prefix sum in chunks of 32
  unsigned int carry = 0;
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int k = k0 + lane;
    unsigned int v = k < n ? (unsigned int)(((unsigned long long)one * s_cdf[k]) / total) : 0u;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    v += carry;
    if (k < n) s_cdf[k] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();
  if (lane == 0) s_cdf[n - 1] = one;
  __syncwarp();
  // give every zero-width symbol one count, stolen from the lowest-frequency symbol with more than one
  for (int i = 0; i < n - 1; ++i) {
    if (s_cdf[i] != s_cdf[i + 1]) continue;  // uniform: every lane reads the same two words
    unsigned int best_f = 0xffffffffu;
    int best_j = 0x7fffffff;
    for (int j = lane; j < n - 1; j += 32) {
      const unsigned int f = s_cdf[j + 1] - s_cdf[j];
      if (f > 1u && f < best_f) { best_f = f; best_j = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned int of = __shfl_xor_sync(0xffffffffu, best_f, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
      if (of < best_f || (of == best_f && oj < best_j)) { best_f = of; best_j = oj; }
    }
    if (best_f == 0xffffffffu) {  // (the reference asserts) nothing to steal from
      if (lane == 0) atomicCAS(status, 0, 1 + row);
      break;
    }
    __syncwarp();
    if (best_j < i) {
      for (int j = best_j + 1 + lane; j <= i; j += 32) s_cdf[j] -= 1u;
    } else {
      for (int j = i + 1 + lane; j <= best_j; j += 32) s_cdf[j] += 1u;
    }
    __syncwarp();
  }
  __syncwarp();
  for (int k = lane; k < cdf_ld; k += 32) cdf_out[(size_t)row * cdf_ld + k] = k < n ? (int)s_cdf[k] : 0;
}

}  // namespace

extern "C" int rdsic_gc_cdf_sizes(const float* table, int32_t rows, float multiplier, int32_t* offset, int32_t* cdf_length,
                                  rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(table && offset && cdf_length && rows > 0 && multiplier > 0.f);
  gc_cdf_sizes_kernel<<<ceil_div(rows, 128), 128, 0, (cudaStream_t)stream>>>(table, rows, multiplier, offset, cdf_length);
  return rdsic_launch_status();
}

extern "C" int rdsic_gc_pmf(const float* table, const int32_t* offset, int32_t rows, float* prob, int32_t ld,
                            rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(table && offset && prob && rows > 0 && ld > 0);
  gc_pmf_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(table, offset, prob, ld);
  return rdsic_launch_status();
}

extern "C" int rdsic_eb_cdf_sizes(const float* quantiles, int32_t C, int32_t* offset, int32_t* cdf_length, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(quantiles && offset && cdf_length && C > 0);
  eb_cdf_sizes_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(quantiles, C, offset, cdf_length);
  return rdsic_launch_status();
}

extern "C" int rdsic_eb_pmf(const float* params, const float* quantiles, const int32_t* offset, const int32_t* cdf_length,
                            int32_t C, int32_t max_length, float* prob, int32_t ld, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(params && quantiles && offset && cdf_length && prob && C > 0 && max_length > 0 && ld > max_length);
  eb_pmf_kernel<<<C, 64, 0, (cudaStream_t)stream>>>(params, quantiles, offset, cdf_length, max_length, prob, ld);
  return rdsic_launch_status();
}

extern "C" int rdsic_pmf_to_quantized_cdf(const float* prob, int32_t ld, const int32_t* cdf_length, int32_t rows,
                                          int32_t precision, int32_t* cdf, int32_t cdf_ld, int32_t* status,
                                          rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(prob && cdf_length && cdf && status && rows > 0 && ld > 0 && cdf_ld >= 2 && cdf_ld <= ld + 1);
  RDSIC_CHECK_ARG(precision >= 1 && precision <= 24);
  const size_t smem = (size_t)cdf_ld * sizeof(unsigned int);  // every row fits cdf_ld entries
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return RDSIC_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(pmf_to_cdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  pmf_to_cdf_kernel<<<rows, 32, smem, (cudaStream_t)stream>>>(prob, ld, cdf_length, precision, cdf, cdf_ld, status);
  return rdsic_launch_status();
}

extern "C" int rdsic_eb_aux_loss(const float* params, const float* quantiles, const float* target, int32_t C, float* terms,
                                 float* sum, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(params && quantiles && target && sum && C > 0);
  eb_aux_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(params, quantiles, target, C, terms, sum);
  return rdsic_launch_status();
}

// ------------------------------------------------------------------ ResDSIC importance mask
// layers/mask_layer.py:64-107 (eval: Mask.forward, then apply_noise's torch.round :37-38), elementwise over
// channels-last fp32 logits.  torch.sigmoid / torch.pow / torch.round -> 1/(1+expf(-v)), powf, rintf.
__global__ void __launch_bounds__(256) mask_forward_kernel(const rdsic_mask_desc d) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)d.B * d.H * d.W * d.C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = e / d.C;
    const int c = (int)(e % d.C);
    float m;
    if (d.mode == 1) {
      const float v = ((const float*)d.in[0].ptr)[pix * d.in[0].ld + d.in[0].coff + c];
      m = powf(__fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))), d.gamma[c]);
    } else {
      float acc = 0.f;
      for (int i = 0; i < d.n_in; ++i) {
        const float v = ((const float*)d.in[i].ptr)[pix * d.in[i].ld + d.in[i].coff + c];
        acc = __fadd_rn(acc, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))));
      }
      m = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-acc)));
    }
    ((float*)d.out.ptr)[pix * d.out.ld + d.out.coff + c] = rintf(m);
  }
}

extern "C" int rdsic_mask_forward(const rdsic_mask_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->out.ptr && d->out.dtype == RDSIC_F32 && !d->out.nchw);
  RDSIC_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->C > 0 && (d->mode == 1 || d->mode == 2));
  RDSIC_CHECK_ARG(d->n_in >= 1 && d->n_in <= 8 && (d->mode != 1 || (d->n_in == 1 && d->gamma)));
  for (int i = 0; i < d->n_in; ++i) RDSIC_CHECK_ARG(d->in[i].ptr && d->in[i].dtype == RDSIC_F32 && !d->in[i].nchw);
  const size_t total = (size_t)d->B * d->H * d->W * d->C;
  const size_t want = (total + 255) / 256, cap = (size_t)rdsic_sm_count() * 8;
  return rdsic_launch(mask_forward_kernel, dim3((unsigned)(want < cap ? want : cap)), 256, 0, (cudaStream_t)stream, false, *d);
}

extern "C" int rdsic_eb_forward(const rdsic_eb_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->z.ptr && d->z_hat.ptr && d->lik && d->params);
  RDSIC_CHECK_ARG(d->B > 0 && d->h > 0 && d->w > 0 && d->C > 0);
  RDSIC_CHECK_ARG(d->z.dtype == RDSIC_F32 && !d->z.nchw && !d->z_hat.nchw);
  RDSIC_CHECK_ARG(!d->noise.ptr || (d->noise.dtype == RDSIC_F32 && !d->noise.nchw));
  RDSIC_CHECK_ARG(!d->noisy_out.ptr || (d->noise.ptr && d->noisy_out.dtype == RDSIC_F32 && !d->noisy_out.nchw));
  const size_t total = (size_t)d->B * d->h * d->w * d->C;
  const size_t smem = (size_t)d->C * EB_PITCH * sizeof(float);
  RDSIC_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(eb_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const size_t want = (total + 255) / 256, cap = (size_t)rdsic_sm_count() * 2;
  const int nblk = (int)(want < cap ? want : cap);
  return rdsic_launch(eb_forward_kernel, dim3((unsigned)nblk), 256, smem, (cudaStream_t)stream, false, *d);
}

extern "C" int rdsic_gc_forward(const rdsic_gc_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && (d->y.ptr || d->sym_in) && d->mu.ptr && d->scale.ptr && d->lik && d->table);
  RDSIC_CHECK_ARG(!(d->sym_in && d->noise.ptr));
  RDSIC_CHECK_ARG(d->B > 0 && d->h > 0 && d->w > 0 && d->Cs > 0 && d->Ctot >= d->lik_coff + d->Cs);
  RDSIC_CHECK_ARG(d->n_table >= 2 && d->n_table <= 129);
  RDSIC_CHECK_ARG((d->sym_in || d->y.dtype == RDSIC_F32) && d->mu.dtype == RDSIC_F32 && d->scale.dtype == RDSIC_F32);
  RDSIC_CHECK_ARG((d->sym_in || !d->y.nchw) && !d->mu.nchw && !d->scale.nchw);
  RDSIC_CHECK_ARG(!d->noise.ptr || (d->noise.dtype == RDSIC_F32 && !d->noise.nchw));
  RDSIC_CHECK_ARG(!d->noisy_out.ptr || (d->noise.ptr && d->noisy_out.dtype == RDSIC_F32 && !d->noisy_out.nchw));
  RDSIC_CHECK_ARG(!d->mask.ptr || (d->mask.dtype == RDSIC_F32 && !d->mask.nchw));
  const size_t npix = (size_t)d->B * d->h * d->w;
  dim3 grid((unsigned)((npix + GC_TP - 1) / GC_TP), (unsigned)ceil_div(d->Cs, GC_TC));
  // 128-bit path: every view addressed in whole float4 / bf16x4 units, NCHW quads inside one image
  auto v4 = [](const rdsic_view& v, int esz) {
    return !v.ptr || (v.ld % 4 == 0 && v.coff % 4 == 0 && ((uintptr_t)v.ptr % (size_t)(4 * esz)) == 0);
  };
  bool vec = d->Cs % 4 == 0 && (d->h * d->w) % 4 == 0 && d->lik_coff >= 0 && v4(d->mu, 4) && v4(d->scale, 4) &&
             (d->sym_in || v4(d->y, 4)) && v4(d->noise, 4) && v4(d->noisy_out, 4) && v4(d->mask, 4) &&
             ((uintptr_t)d->lik % 16) == 0 && ((uintptr_t)d->symbols % 16) == 0 && ((uintptr_t)d->indexes % 16) == 0;
  for (int k = 0; k < 3; ++k) vec = vec && v4(d->y_hat[k], d->y_hat[k].dtype == RDSIC_BF16 ? 2 : 4);
  static const int tune_vec = getenv("RDSIC_GC_VEC") ? atoi(getenv("RDSIC_GC_VEC")) : 1;
  if (vec && tune_vec) return rdsic_launch(gc_forward_vec_kernel, grid, 256, 0, (cudaStream_t)stream, false, *d);
  return rdsic_launch(gc_forward_kernel, grid, 256, 0, (cudaStream_t)stream, false, *d);
}
