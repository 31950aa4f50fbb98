// Backward ("_bwd") kernels of the rate-distortion training step (BASELINE config 4), fp32.
//
// The reference trains through torch autograd (training/step.py:42-56: forward, RateDistortionLoss
// training/loss.py:6-30, backward, aux_loss backward, clip, two Adam steps).  Here every node of that graph whose
// forward is one of this library's kernels gets a hand-written backward twin; resdsic_b200/training/functions.py
// wraps the pairs as torch.autograd.Function nodes (torch supplies the graph walk and the optimisers, no arithmetic):
//
//   rdsic_conv_dgrad_f32     d input of conv / linear / GDN contraction (implicit GEMM over the transposed taps)
//   rdsic_conv_wgrad_f32     d weight (+ d bias): reduction over pixels, packed [Cout][KH*KW*Cin] like the forward
//   rdsic_pointwise_f32      forward AND backward of the elementwise nodes (GELU, sigmoid gate, GDN scaling, square,
//                            LRP tanh, adds, loss scalings) -- in training mode they are separate nodes because the
//                            backward needs the pre-activation values the fused inference epilogues never store
//   rdsic_pixel_shuffle_f32  nn.PixelShuffle(2) on NHWC and its inverse (subpel_conv3x3, layers/layers.py:34-38)
//   rdsic_attn_backward_f32  window attention core: d qkv and d relative_position_bias_table
//   rdsic_gc_backward        GaussianConditional likelihood (noise mode) + ste_round: d y, d mu, d scale with the
//                            LowerBound gradient rule (ops/bound_ops.py:21-27) on scale (0.11) and likelihood (1e-9)
//   rdsic_eb_backward        EntropyBottleneck likelihood (noise mode): d z and d of the packed per-channel
//                            parameters (softplus(matrix), bias, tanh(factor)); sign detached (entropy_models.py:429-430)
//   rdsic_eb_aux_backward    aux_loss: d quantiles only (entropy_models.py:396-399)
//   rdsic_reduce_f32         sum(log x) and sum((a-b)^2) in fp64 (RateDistortionLoss)
#include <stdlib.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ conv dgrad
// dx[b,iy,ix,ci] = sum_{r,s,co} dpre[b,oy,ox,co] * W[co][r][s][ci]  with  iy = oy*stride - pad + r  (same for x).
// Implicit GEMM: M = B*H*W input pixels, N = Cin, K = taps*Cout; the weight is packed [Cin][taps*Cout].
constexpr int DBM = 128, DBN = 64, DBK = 16, DNT = 256;

__global__ void __launch_bounds__(DNT) conv_dgrad_f32_kernel(const rdsic_conv_desc d, const float* __restrict__ wt) {
  __shared__ __align__(16) float As[DBK][DBM + 4];
  __shared__ __align__(16) float Bs[DBK][DBN + 4];
  const int M = d.B * d.H * d.W;
  const int K = d.KH * d.KW * d.Cout;
  const int m0 = blockIdx.x * DBM, n0 = blockIdx.y * DBN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const float* __restrict__ dp = (const float*)d.out.ptr;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ar = tid / 4, ak4 = tid % 4;
  int rb[2], ry[2], rx[2];
  bool rv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ar + 64 * i;
    rv[i] = m < M;
    const int mm = rv[i] ? m : 0;
    rx[i] = mm % d.W + d.pad_w;
    const int t = mm / d.W;
    ry[i] = t % d.H + d.pad_h;
    rb[i] = t / d.H;
  }
  float4 ra[2], rbv;
  auto load_tile = [&](int k0) {
    const int tap = k0 / d.Cout, c0 = k0 - tap * d.Cout;
    const int r = tap / d.KW, s = tap - r * d.KW;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ty_ = ry[i] - r, tx_ = rx[i] - s;
      if (rv[i] && ty_ >= 0 && tx_ >= 0 && ty_ % d.stride == 0 && tx_ % d.stride == 0) {
        const int oy = ty_ / d.stride, ox = tx_ / d.stride;
        if (oy < d.OH && ox < d.OW) {
          const size_t pix = ((size_t)rb[i] * d.OH + oy) * d.OW + ox;
          v = *reinterpret_cast<const float4*>(dp + pix * d.out.ld + d.out.coff + c0 + ak4 * 4);
        }
      }
      ra[i] = v;
    }
    const int n = n0 + ar;
    rbv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < d.Cin) rbv = *reinterpret_cast<const float4*>(wt + (size_t)n * K + k0 + ak4 * 4);
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = ar + i * 64;
      As[ak4 * 4 + 0][row] = ra[i].x;
      As[ak4 * 4 + 1][row] = ra[i].y;
      As[ak4 * 4 + 2][row] = ra[i].z;
      As[ak4 * 4 + 3][row] = ra[i].w;
    }
    Bs[ak4 * 4 + 0][ar] = rbv.x;
    Bs[ak4 * 4 + 1][ar] = rbv.y;
    Bs[ak4 * 4 + 2][ar] = rbv.z;
    Bs[ak4 * 4 + 3][ar] = rbv.w;
  };
  const int nk = K / DBK;
  load_tile(0);
  store_tile();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tile((kt + 1) * DBK);
#pragma unroll
    for (int kk = 0; kk < DBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (kt + 1 < nk) {
      store_tile();
      __syncthreads();
    }
  }
  float* __restrict__ dx = (float*)d.in.ptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d.Cin) dx[(size_t)m * d.in.ld + d.in.coff + n] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------ conv wgrad
// dW[co][tap*Cin + ci] = sum_m dpre[m][co] * x[b, oy*stride - pad + r, ox*stride - pad + s, ci];  db[co] = sum_m dpre[m][co].
// Tile 64 (co) x 64 (k), the pixels split over gridDim.z; partial sums meet in fp32 atomics (dW, db zeroed first).
constexpr int WBM = 64, WBN = 64, WBP = 16, WNT = 256;

template <bool FAST>
__global__ void __launch_bounds__(WNT) conv_wgrad_f32_kernel(const rdsic_conv_desc d, float* __restrict__ dw,
                                                           float* __restrict__ db, int pix_per_split) {
  __shared__ __align__(16) float As[WBP][WBM + 4];  // dpre  [pixel][co]
  __shared__ __align__(16) float Bs[WBP][WBN + 4];  // patch [pixel][k]
  const int M = d.B * d.OH * d.OW;
  const int K = d.KH * d.KW * d.Cin;
  const int co0 = blockIdx.x * WBM, k0 = blockIdx.y * WBN;
  const int p_begin = blockIdx.z * pix_per_split;
  const int p_end = min(M, p_begin + pix_per_split);
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const float* __restrict__ dp = (const float*)d.out.ptr;
  const float* __restrict__ in = (const float*)d.in.ptr;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};

  // loader: thread -> (pixel lp, 4 consecutive columns lc4) of both tiles
  const int lp = tid / 16, lc4 = tid % 16;
  // FAST: the 4 columns k0 + 4 lc4 .. +3 of the patch tile are 4 consecutive channels of ONE tap (Cin % 4 == 0)
  int f_r = 0, f_s = 0, f_c = 0;
  bool f_ok = false;
  if (FAST) {
    const int k = k0 + 4 * lc4;
    f_ok = k < K;
    const int kk = f_ok ? k : 0;
    const int tap = kk / d.Cin;
    f_c = kk - tap * d.Cin;
    f_r = tap / d.KW;
    f_s = tap - f_r * d.KW;
  }
  for (int p0 = p_begin; p0 < p_end; p0 += WBP) {
    const int m = p0 + lp;
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (m < p_end) {
      const int ox = m % d.OW;
      const int t = m / d.OW;
      const int oy = t % d.OH, b = t / d.OH;
      const int co = co0 + 4 * lc4;
      const float* dpp = dp + (size_t)m * d.out.ld + d.out.coff;
      if (co + 3 < d.Cout && (d.out.ld % 4 == 0) && (d.out.coff % 4 == 0)) {
        va = *reinterpret_cast<const float4*>(dpp + co);
      } else {
        if (co < d.Cout) va.x = dpp[co];
        if (co + 1 < d.Cout) va.y = dpp[co + 1];
        if (co + 2 < d.Cout) va.z = dpp[co + 2];
        if (co + 3 < d.Cout) va.w = dpp[co + 3];
      }
      if (FAST) {
        const int iy = oy * d.stride - d.pad_h + f_r, ix = ox * d.stride - d.pad_w + f_s;
        if (f_ok && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W)
          vb = *reinterpret_cast<const float4*>(in + (((size_t)b * d.H + iy) * d.W + ix) * d.in.ld + d.in.coff + f_c);
      } else {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + 4 * lc4 + q;
          if (k >= K) continue;
          const int tap = k / d.Cin, c = k - tap * d.Cin;
          const int r = tap / d.KW, s = tap - r * d.KW;
          const int iy = oy * d.stride - d.pad_h + r, ix = ox * d.stride - d.pad_w + s;
          if (iy < 0 || iy >= d.H || ix < 0 || ix >= d.W) continue;
          e[q] = d.in.nchw ? in[(((size_t)b * d.Cin + c) * d.H + iy) * d.W + ix]
                           : in[(((size_t)b * d.H + iy) * d.W + ix) * d.in.ld + d.in.coff + c];
        }
        vb = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
    __syncthreads();  // previous tile fully consumed
    *reinterpret_cast<float4*>(&As[lp][4 * lc4]) = va;
    *reinterpret_cast<float4*>(&Bs[lp][4 * lc4]) = vb;
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < WBP; ++pp) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[pp][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[pp][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= d.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dw + (size_t)co * K + k, acc[i][j]);
    }
    if (db && blockIdx.y == 0 && tx == 0) atomicAdd(db + co, bsum[i]);
  }
}

// ------------------------------------------------------------------ pointwise forward / backward nodes
__device__ __forceinline__ float gelu_grad(float x) {  // d/dx 0.5 x (1 + erf(x / sqrt 2))
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void pointwise_f32_kernel(int op, size_t n, const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ c, float* __restrict__ o0, float* __restrict__ o1, float alpha) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    switch (op) {
      case RDSIC_PW_ADD: o0[i] = a[i] + b[i]; break;
      case RDSIC_PW_GELU_FWD: o0[i] = gelu_erf(a[i]); break;
      case RDSIC_PW_GELU_BWD: o0[i] = a[i] * gelu_grad(b[i]); break;
      case RDSIC_PW_GATE_FWD: o0[i] = a[i] * sigmoid_f(b[i]) + c[i]; break;
      case RDSIC_PW_GATE_BWD: {  // a = g, b = gate operand a, c = gate logits
        const float s = sigmoid_f(c[i]);
        o0[i] = a[i] * s;
        o1[i] = a[i] * b[i] * s * (1.0f - s);
        break;
      }
      case RDSIC_PW_GDN_FWD: o0[i] = alpha > 0.f ? a[i] * sqrtf(b[i]) : a[i] * rsqrtf(b[i]); break;
      case RDSIC_PW_GDN_BWD: {  // a = g, b = x, c = norm;  o0 = d x (direct), o1 = d norm
        const float nn = c[i];
        if (alpha > 0.f) {
          const float sq = sqrtf(nn);
          o0[i] = a[i] * sq;
          o1[i] = a[i] * b[i] * 0.5f / sq;
        } else {
          const float rs = rsqrtf(nn);
          o0[i] = a[i] * rs;
          o1[i] = a[i] * b[i] * (-0.5f) * rs / nn;
        }
        break;
      }
      case RDSIC_PW_SQUARE_FWD: o0[i] = a[i] * a[i]; break;
      case RDSIC_PW_SQUARE_BWD: o0[i] = 2.0f * a[i] * b[i]; break;
      case RDSIC_PW_LRP_FWD: o0[i] = a[i] + 0.5f * tanhf(b[i]); break;
      case RDSIC_PW_LRP_BWD: {
        const float t = tanhf(b[i]);
        o0[i] = a[i] * 0.5f * (1.0f - t * t);
        break;
      }
      case RDSIC_PW_RECIP_SCALE: o0[i] = alpha / a[i]; break;
      case RDSIC_PW_DIFF_SCALE: o0[i] = alpha * (a[i] - b[i]); break;
      case RDSIC_PW_SCALE: o0[i] = alpha * a[i]; break;
      case RDSIC_PW_MUL: o0[i] = a[i] * b[i]; break;
      default: break;
    }
  }
}

// NHWC PixelShuffle(2): out[b, 2y+i, 2x+j, c] = in[b, y, x, 4c + 2i + j]; dir 1 = inverse.  (H, W, C) are the INPUT's
// spatial size and the OUTPUT channel count of the forward direction.
__global__ void pixel_shuffle_f32_kernel(int dir, const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C) {
  const size_t total = (size_t)B * H * W * C * 4;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    // e indexes the shuffled tensor [B, 2H, 2W, C]
    const int c = (int)(e % C);
    size_t t = e / C;
    const int X = (int)(t % (2 * W));
    t /= 2 * W;
    const int Y = (int)(t % (2 * H));
    const int b = (int)(t / (2 * H));
    const size_t src = ((((size_t)b * H + (Y >> 1)) * W + (X >> 1)) * (size_t)(4 * C)) + 4 * c + 2 * (Y & 1) + (X & 1);
    if (dir == 0) out[e] = in[src];
    else out[src] = in[e];
  }
}

// ------------------------------------------------------------------ window attention backward
// One CTA per (window, head), thread i = query token i (as win_attn_f32_kernel).  P and dS are staged in shared
// memory so that thread j can form the column sums dK_j, dV_j.
template <int NTOK>
__global__ void __launch_bounds__(NTOK) win_attn_bwd_f32_kernel(const rdsic_attn_desc d, const rdsic_view dout,
                                                              const rdsic_view dqkv, float* __restrict__ dbias) {
  extern __shared__ float smem[];
  const int ws = d.ws, C = d.C, heads = d.heads, dh = C / heads;
  const int ldq = dh + 1;
  float* qs = smem;                    // [NTOK][dh+1]  q * scale
  float* ks = qs + NTOK * ldq;         // [NTOK][dh]
  float* vs = ks + NTOK * dh;          // [NTOK][dh]
  float* gs = vs + NTOK * dh;          // [NTOK][dh+1]  dO
  float* Ps = gs + NTOK * ldq;         // [NTOK][NTOK+1]
  float* Ss = Ps + NTOK * (NTOK + 1);  // [NTOK][NTOK+1] dS
  int* rid = (int*)(Ss + NTOK * (NTOK + 1));

  const int head = blockIdx.x % heads;
  int win = blockIdx.x / heads;
  const int nWw = d.W / ws, nWh = d.H / ws;
  const int ww = win % nWw;
  win /= nWw;
  const int wh = win % nWh;
  const int b = win / nWh;
  const float scale = d.scale;

  const int i = threadIdx.x;
  const int hy = wh * ws + i / ws, wx = ww * ws + i % ws;
  const int oy = (hy + d.shift) % d.H, ox = (wx + d.shift) % d.W;
  const size_t pix = ((size_t)b * d.H + oy) * d.W + ox;
  {
    const int rh = (hy >= d.H - ws) + (hy >= d.H - d.shift), rw = (wx >= d.W - ws) + (wx >= d.W - d.shift);
    rid[i] = d.shift > 0 ? 3 * rh + rw : 0;
  }
  const float* qkv = (const float*)d.qkv.ptr + pix * d.qkv.ld + d.qkv.coff + head * dh;
  const float* go = (const float*)dout.ptr + pix * dout.ld + dout.coff + head * dh;
  for (int c = 0; c < dh; ++c) {
    qs[i * ldq + c] = qkv[c] * scale;
    ks[i * dh + c] = qkv[C + c];
    vs[i * dh + c] = qkv[2 * C + c];
    gs[i * ldq + c] = go[c];
  }
  __syncthreads();

  const int hi = i / ws, wi = i % ws, myrid = rid[i];
  const int tw = 2 * ws - 1;
  float* Pi = Ps + i * (NTOK + 1);
  float* Si = Ss + i * (NTOK + 1);
  float mx = -INFINITY;
  for (int j = 0; j < NTOK; ++j) {
    float a = 0.f;
    for (int c = 0; c < dh; ++c) a = fmaf(qs[i * ldq + c], ks[j * dh + c], a);
    a += d.bias_table[((hi - j / ws + ws - 1) * tw + (wi - j % ws + ws - 1)) * heads + head];
    if (rid[j] != myrid) a += -100.0f;
    Pi[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
  for (int j = 0; j < NTOK; ++j) {
    const float e = expf(Pi[j] - mx);
    Pi[j] = e;
    sum += e;
  }
  // dP_ij = dO_i . V_j ;  dS_ij = P_ij (dP_ij - sum_j' P_ij' dP_ij')
  float dot = 0.f;
  for (int j = 0; j < NTOK; ++j) {
    const float p = Pi[j] / sum;
    Pi[j] = p;
    float dpv = 0.f;
    for (int c = 0; c < dh; ++c) dpv = fmaf(gs[i * ldq + c], vs[j * dh + c], dpv);
    Si[j] = dpv;
    dot = fmaf(p, dpv, dot);
  }
  for (int j = 0; j < NTOK; ++j) {
    const float ds = Pi[j] * (Si[j] - dot);
    Si[j] = ds;
    atomicAdd(dbias + ((hi - j / ws + ws - 1) * tw + (wi - j % ws + ws - 1)) * heads + head, ds);
  }
  float* gq = (float*)dqkv.ptr + pix * dqkv.ld + dqkv.coff + head * dh;
  // dq_i = scale * sum_j dS_ij K_j
  for (int c = 0; c < dh; ++c) {
    float a = 0.f;
    for (int j = 0; j < NTOK; ++j) a = fmaf(Si[j], ks[j * dh + c], a);
    gq[c] = a * scale;
  }
  __syncthreads();
  // thread i now acts as key / value token i:  dK_i = sum_q dS_qi (q_q * scale) ;  dV_i = sum_q P_qi dO_q
  for (int c = 0; c < dh; ++c) {
    float ak = 0.f, av = 0.f;
    for (int q = 0; q < NTOK; ++q) {
      ak = fmaf(Ss[q * (NTOK + 1) + i], qs[q * ldq + c], ak);
      av = fmaf(Ps[q * (NTOK + 1) + i], gs[q * ldq + c], av);
    }
    gq[C + c] = ak;
    gq[2 * C + c] = av;
  }
}

// ------------------------------------------------------------------ GaussianConditional backward (noise mode)
// forward (entropy_models.py:627-661, training): t = y + noise - mu, v = |t|, s = max(scale, 0.11),
//   lik = max(Phi((.5 - v)/s) - Phi((-.5 - v)/s), 1e-9);  y_hat = ste_round(y - mu) + mu (cnn.py:177).
// LowerBound rule (bound_ops.py:25-27): the gradient passes where x >= bound or where it is negative.
__global__ void gc_backward_kernel(const rdsic_gc_desc d, const float* __restrict__ g_lik, const rdsic_view g_yhat,
                                   const rdsic_view dy, const rdsic_view dmu, const rdsic_view dscale) {
  const int hw = d.h * d.w;
  const size_t total = (size_t)d.B * hw * d.Cs;
  const float inv_sqrt2pi = 0.39894228040143267794f;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % d.Cs);
    const size_t pix = e / d.Cs;
    const float y = ((const float*)d.y.ptr)[pix * d.y.ld + d.y.coff + c];
    const float mu = ((const float*)d.mu.ptr)[pix * d.mu.ld + d.mu.coff + c];
    const float sc = ((const float*)d.scale.ptr)[pix * d.scale.ld + d.scale.coff + c];
    const float nz = d.noise.ptr ? ((const float*)d.noise.ptr)[pix * d.noise.ld + d.noise.coff + c] : 0.f;
    const size_t b = pix / hw, yx = pix % hw;
    const float g = g_lik[(b * d.Ctot + d.lik_coff + c) * (size_t)hw + yx];
    const float t = (y + nz) - mu;
    const float v = fabsf(t);
    const float s = fmaxf(sc, d.scale_bound);
    const float u = (0.5f - v) / s, l = (-0.5f - v) / s;
    const float cst = -0.70710678118654752440f;
    const float lik = 0.5f * erfcf(cst * u) - 0.5f * erfcf(cst * l);
    const float g1 = (lik >= d.lik_bound || g < 0.f) ? g : 0.f;  // likelihood_lower_bound
    const float pu = inv_sqrt2pi * expf(-0.5f * u * u), pl = inv_sqrt2pi * expf(-0.5f * l * l);
    const float dv = g1 * (pl - pu) / s;
    float ds = g1 * (l * pl - u * pu) / s;
    ds = (sc >= d.scale_bound || ds < 0.f) ? ds : 0.f;  // lower_bound_scale
    const float sg = t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f);
    const float dt = dv * sg;
    float gy = dt;
    if (g_yhat.ptr) gy += ((const float*)g_yhat.ptr)[pix * g_yhat.ld + g_yhat.coff + c];  // ste_round(y - mu) + mu: d/dy = 1, d/dmu = 0
    ((float*)dy.ptr)[pix * dy.ld + dy.coff + c] = gy;
    ((float*)dmu.ptr)[pix * dmu.ld + dmu.coff + c] = -dt;
    ((float*)dscale.ptr)[pix * dscale.ld + dscale.coff + c] = ds;
  }
}

// ------------------------------------------------------------------ EntropyBottleneck backward (noise mode)
constexpr int EB_M0 = 0, EB_M1 = 3, EB_M2 = 12, EB_M3 = 21, EB_M4 = 30;
constexpr int EB_B0 = 33, EB_B1 = 36, EB_B2 = 39, EB_B3 = 42, EB_B4 = 45;
constexpr int EB_F0 = 46, EB_F1 = 49, EB_F2 = 52, EB_F3 = 55;
constexpr int EB_NP = 58;  // parameters with a gradient (entry 58 is the median)

// forward of the cumulative-logits chain at v, keeping the four hidden pre-activations; returns the logit
__device__ __forceinline__ float eb_chain_fwd(const float* p, float v, float pre[4][3], float h[4][3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pre[0][j] = p[EB_M0 + j] * v + p[EB_B0 + j];
    h[0][j] = pre[0][j] + p[EB_F0 + j] * tanhf(pre[0][j]);
  }
  const int mo[3] = {EB_M1, EB_M2, EB_M3}, bo[3] = {EB_B1, EB_B2, EB_B3}, fo[3] = {EB_F1, EB_F2, EB_F3};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* m = p + mo[k] + 3 * j;
      pre[k + 1][j] = m[0] * h[k][0] + m[1] * h[k][1] + m[2] * h[k][2] + p[bo[k] + j];
      h[k + 1][j] = pre[k + 1][j] + p[fo[k] + j] * tanhf(pre[k + 1][j]);
    }
  return p[EB_M4] * h[3][0] + p[EB_M4 + 1] * h[3][1] + p[EB_M4 + 2] * h[3][2] + p[EB_B4];
}
// backward of the chain: accumulates g * d logit / d param into gp[EB_NP] (if gp) and returns g * d logit / d v
__device__ __forceinline__ float eb_chain_bwd(const float* p, float v, float g, const float pre[4][3], const float h[4][3], float* gp) {
  float dh[3], dpre[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (gp) gp[EB_M4 + j] += g * h[3][j];
    dh[j] = g * p[EB_M4 + j];
  }
  if (gp) gp[EB_B4] += g;
  const int mo[3] = {EB_M1, EB_M2, EB_M3}, bo[3] = {EB_B1, EB_B2, EB_B3}, fo[3] = {EB_F1, EB_F2, EB_F3};
#pragma unroll
  for (int k = 2; k >= 0; --k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float t = tanhf(pre[k + 1][j]);
      if (gp) gp[fo[k] + j] += dh[j] * t;
      dpre[j] = dh[j] * (1.0f + p[fo[k] + j] * (1.0f - t * t));
      if (gp) gp[bo[k] + j] += dpre[j];
    }
    float nh[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (gp) gp[mo[k] + 3 * j + q] += dpre[j] * h[k][q];
        nh[q] += p[mo[k] + 3 * j + q] * dpre[j];
      }
#pragma unroll
    for (int q = 0; q < 3; ++q) dh[q] = nh[q];
  }
  float dv = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float t = tanhf(pre[0][j]);
    if (gp) gp[EB_F0 + j] += dh[j] * t;
    const float dp0 = dh[j] * (1.0f + p[EB_F0 + j] * (1.0f - t * t));
    if (gp) {
      gp[EB_B0 + j] += dp0;
      gp[EB_M0 + j] += dp0 * v;
    }
    dv += dp0 * p[EB_M0 + j];
  }
  return dv;
}

// one block per channel: elements strided over the block, per-thread parameter gradients, block tree reduction
constexpr int EBB_NT = 128;
__global__ void __launch_bounds__(EBB_NT) eb_backward_kernel(const rdsic_eb_desc d, const float* __restrict__ g_lik,
                                                           const rdsic_view g_zhat, const rdsic_view dz, float* __restrict__ dparams) {
  __shared__ float red[EBB_NT];
  const int c = blockIdx.x;
  const int hw = d.h * d.w;
  const size_t npix = (size_t)d.B * hw;
  float p[RDSIC_EB_STRIDE];
  for (int k = 0; k < RDSIC_EB_STRIDE; ++k) p[k] = d.params[(size_t)c * RDSIC_EB_STRIDE + k];
  float gp[EB_NP];
  for (int k = 0; k < EB_NP; ++k) gp[k] = 0.f;
  for (size_t pix = threadIdx.x; pix < npix; pix += EBB_NT) {
    const float z = ((const float*)d.z.ptr)[pix * d.z.ld + d.z.coff + c];
    const float nz = d.noise.ptr ? ((const float*)d.noise.ptr)[pix * d.noise.ld + d.noise.coff + c] : 0.f;
    const float v = z + nz;
    const size_t b = pix / hw, yx = pix % hw;
    const float g = g_lik[(b * d.C + c) * (size_t)hw + yx];
    float prl[4][3], hl[4][3], pru[4][3], hu[4][3];
    const float lo = eb_chain_fwd(p, v - 0.5f, prl, hl), up = eb_chain_fwd(p, v + 0.5f, pru, hu);
    const float su = lo + up;
    const float sgn = su > 0.f ? -1.f : (su < 0.f ? 1.f : 0.f);  // detached (entropy_models.py:429-430)
    const float s_up = sigmoid_f(sgn * up), s_lo = sigmoid_f(sgn * lo);
    const float diff = s_up - s_lo;
    const float lik = fabsf(diff);
    const float g1 = (lik >= d.lik_bound || g < 0.f) ? g : 0.f;
    const float sd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const float g_up = g1 * sd * s_up * (1.0f - s_up) * sgn;
    const float g_lo = -g1 * sd * s_lo * (1.0f - s_lo) * sgn;
    float gz = eb_chain_bwd(p, v + 0.5f, g_up, pru, hu, gp) + eb_chain_bwd(p, v - 0.5f, g_lo, prl, hl, gp);
    if (g_zhat.ptr) gz += ((const float*)g_zhat.ptr)[pix * g_zhat.ld + g_zhat.coff + c];  // z_hat = ste_round(z - med) + med
    ((float*)dz.ptr)[pix * dz.ld + dz.coff + c] = gz;
  }
  for (int k = 0; k < EB_NP; ++k) {
    red[threadIdx.x] = gp[k];
    __syncthreads();
    for (int w = EBB_NT / 2; w > 0; w >>= 1) {
      if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) dparams[(size_t)c * RDSIC_EB_STRIDE + k] = red[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dparams[(size_t)c * RDSIC_EB_STRIDE + 58] = 0.f;
    dparams[(size_t)c * RDSIC_EB_STRIDE + 59] = 0.f;
  }
}

// aux_loss = sum_{c,k} |logits_c(quantiles[c,k]) - target[k]| with every parameter but `quantiles` detached
__global__ void eb_aux_backward_kernel(const float* __restrict__ params, const float* __restrict__ quantiles,
                                       const float* __restrict__ target, int C, float g, float* __restrict__ dq) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C * 3) return;
  const int c = e / 3, k = e % 3;
  float p[RDSIC_EB_STRIDE];
  for (int q = 0; q < RDSIC_EB_STRIDE; ++q) p[q] = params[(size_t)c * RDSIC_EB_STRIDE + q];
  float pre[4][3], h[4][3];
  const float v = quantiles[e];
  const float lg = eb_chain_fwd(p, v, pre, h);
  const float df = lg - target[k];
  const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
  dq[e] = eb_chain_bwd(p, v, g * sg, pre, h, nullptr);
}

// ------------------------------------------------------------------ reductions (fp64 accumulation)
__global__ void __launch_bounds__(256) reduce_f32_kernel(int op, size_t n, const float* __restrict__ a, const float* __restrict__ b,
                                                        double* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (op == RDSIC_RED_SUM_LOG) acc += (double)logf(a[i]);
    else if (op == RDSIC_RED_SSE) { const double df = (double)a[i] - (double)b[i]; acc += df * df; }
    else acc += (double)a[i];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

int flat_grid(size_t n, int threads) {
  const size_t want = (n + threads - 1) / threads, cap = (size_t)rdsic_sm_count() * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

extern "C" int rdsic_conv_dgrad_f32(const rdsic_conv_desc* d, const float* wt_dgrad, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->in.ptr && d->out.ptr && wt_dgrad);
  RDSIC_CHECK_ARG(d->in.dtype == RDSIC_F32 && d->out.dtype == RDSIC_F32 && !d->in.nchw && !d->out.nchw);
  RDSIC_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->Cout % 16 == 0);
  RDSIC_CHECK_ARG(d->KH > 0 && d->KW > 0 && d->stride > 0 && d->OH > 0 && d->OW > 0);
  RDSIC_CHECK_ARG(d->osy == 1 && d->osx == 1 && d->ooy == 0 && d->oox == 0 && !d->pixel_shuffle);
  if (d->out.ld % 4 || d->out.coff % 4 || ((uintptr_t)d->out.ptr % 16) || ((uintptr_t)wt_dgrad % 16)) return RDSIC_E_ALIGN;
  const int M = d->B * d->H * d->W;
  dim3 grid(ceil_div(M, DBM), ceil_div(d->Cin, DBN));
  conv_dgrad_f32_kernel<<<grid, DNT, 0, (cudaStream_t)stream>>>(*d, wt_dgrad);
  return rdsic_launch_status();
}

extern "C" int rdsic_conv_wgrad_f32(const rdsic_conv_desc* d, float* dw, float* db, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->in.ptr && d->out.ptr && dw);
  RDSIC_CHECK_ARG(d->in.dtype == RDSIC_F32 && d->out.dtype == RDSIC_F32 && !d->out.nchw);
  RDSIC_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0);
  RDSIC_CHECK_ARG(d->KH > 0 && d->KW > 0 && d->stride > 0 && d->OH > 0 && d->OW > 0);
  RDSIC_CHECK_ARG(d->osy == 1 && d->osx == 1 && d->ooy == 0 && d->oox == 0 && !d->pixel_shuffle);
  const int M = d->B * d->OH * d->OW, K = d->KH * d->KW * d->Cin;
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)d->Cout * K * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (db) {
    e = cudaMemsetAsync(db, 0, (size_t)d->Cout * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  const int gx = ceil_div(d->Cout, WBM), gy = ceil_div(K, WBN);
  int splits = ceil_div(4 * rdsic_sm_count(), gx * gy);
  const int max_splits = ceil_div(M, 4 * WBP);
  if (splits > max_splits) splits = max_splits;
  if (splits > 512) splits = 512;
  if (splits < 1) splits = 1;
  const int pps = ceil_div(ceil_div(M, splits), WBP) * WBP;
  splits = ceil_div(M, pps);
  dim3 grid(gx, gy, splits);
  const bool fast = !d->in.nchw && d->Cin % 4 == 0 && d->in.ld % 4 == 0 && d->in.coff % 4 == 0 && ((uintptr_t)d->in.ptr % 16) == 0;
  if (fast) conv_wgrad_f32_kernel<true><<<grid, WNT, 0, (cudaStream_t)stream>>>(*d, dw, db, pps);
  else conv_wgrad_f32_kernel<false><<<grid, WNT, 0, (cudaStream_t)stream>>>(*d, dw, db, pps);
  return rdsic_launch_status();
}

extern "C" int rdsic_pointwise_f32(int op, size_t n, const float* a, const float* b, const float* c, float* o0, float* o1,
                                   float alpha, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(op >= RDSIC_PW_ADD && op <= RDSIC_PW_MUL && a && o0);
  if (n == 0) return 0;
  pointwise_f32_kernel<<<flat_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(op, n, a, b, c, o0, o1, alpha);
  return rdsic_launch_status();
}

extern "C" int rdsic_pixel_shuffle_f32(int inverse, const float* in, float* out, int B, int H, int W, int C, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0);
  const size_t n = (size_t)B * H * W * C * 4;
  pixel_shuffle_f32_kernel<<<flat_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(inverse ? 1 : 0, in, out, B, H, W, C);
  return rdsic_launch_status();
}

extern "C" int rdsic_attn_backward_f32(const rdsic_attn_desc* d, const rdsic_view* dout, const rdsic_view* dqkv, float* dbias,
                                       rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->qkv.ptr && d->bias_table && dout && dout->ptr && dqkv && dqkv->ptr && dbias);
  RDSIC_CHECK_ARG(d->B > 0 && d->heads > 0 && d->C % d->heads == 0 && d->ws > 0 && d->H % d->ws == 0 && d->W % d->ws == 0);
  RDSIC_CHECK_ARG(d->shift >= 0 && d->shift < d->ws);
  RDSIC_CHECK_ARG(d->qkv.dtype == RDSIC_F32 && dout->dtype == RDSIC_F32 && dqkv->dtype == RDSIC_F32);
  RDSIC_CHECK_ARG(!d->qkv.nchw && !dout->nchw && !dqkv->nchw);
  const int ntok = d->ws * d->ws, dh = d->C / d->heads, tw = 2 * d->ws - 1;
  cudaError_t e = cudaMemsetAsync(dbias, 0, (size_t)tw * tw * d->heads * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  const int nblk = d->B * (d->H / d->ws) * (d->W / d->ws) * d->heads;
  const size_t smem = (size_t)(2 * ntok * (dh + 1) + 2 * ntok * dh + 2 * ntok * (ntok + 1)) * sizeof(float) + ntok * sizeof(int);
  if (ntok == 64) {
    auto kern = win_attn_bwd_f32_kernel<64>;
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    kern<<<nblk, 64, smem, (cudaStream_t)stream>>>(*d, *dout, *dqkv, dbias);
  } else if (ntok == 16) {
    win_attn_bwd_f32_kernel<16><<<nblk, 16, smem, (cudaStream_t)stream>>>(*d, *dout, *dqkv, dbias);
  } else {
    return RDSIC_E_UNSUPPORTED;
  }
  return rdsic_launch_status();
}

extern "C" int rdsic_gc_backward(const rdsic_gc_desc* d, const float* g_lik, const rdsic_view* g_yhat, const rdsic_view* dy,
                                 const rdsic_view* dmu, const rdsic_view* dscale, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->y.ptr && d->mu.ptr && d->scale.ptr && g_lik && dy && dmu && dscale && dy->ptr && dmu->ptr && dscale->ptr);
  RDSIC_CHECK_ARG(d->B > 0 && d->h > 0 && d->w > 0 && d->Cs > 0 && d->Ctot >= d->lik_coff + d->Cs);
  RDSIC_CHECK_ARG(d->y.dtype == RDSIC_F32 && d->mu.dtype == RDSIC_F32 && d->scale.dtype == RDSIC_F32);
  RDSIC_CHECK_ARG(!d->y.nchw && !d->mu.nchw && !d->scale.nchw && !d->mask.ptr && !d->sym_in);
  rdsic_view gy = {};
  if (g_yhat) gy = *g_yhat;
  const size_t n = (size_t)d->B * d->h * d->w * d->Cs;
  gc_backward_kernel<<<flat_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(*d, g_lik, gy, *dy, *dmu, *dscale);
  return rdsic_launch_status();
}

extern "C" int rdsic_eb_backward(const rdsic_eb_desc* d, const float* g_lik, const rdsic_view* g_zhat, const rdsic_view* dz,
                                 float* dparams, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->z.ptr && d->params && g_lik && dz && dz->ptr && dparams);
  RDSIC_CHECK_ARG(d->B > 0 && d->h > 0 && d->w > 0 && d->C > 0 && d->z.dtype == RDSIC_F32 && !d->z.nchw);
  rdsic_view gz = {};
  if (g_zhat) gz = *g_zhat;
  eb_backward_kernel<<<d->C, EBB_NT, 0, (cudaStream_t)stream>>>(*d, g_lik, gz, *dz, dparams);
  return rdsic_launch_status();
}

extern "C" int rdsic_eb_aux_backward(const float* params, const float* quantiles, const float* target, int C, float g,
                                     float* dquantiles, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(params && quantiles && target && dquantiles && C > 0);
  eb_aux_backward_kernel<<<ceil_div(C * 3, 128), 128, 0, (cudaStream_t)stream>>>(params, quantiles, target, C, g, dquantiles);
  return rdsic_launch_status();
}

extern "C" int rdsic_reduce_f32(int op, size_t n, const float* a, const float* b, double* out, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(op >= RDSIC_RED_SUM && op <= RDSIC_RED_SSE && a && out && (op != RDSIC_RED_SSE || b));
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return 0;
  int grid = flat_grid(n, 256);
  if (grid > 1024) grid = 1024;
  reduce_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(op, n, a, b, out);
  return rdsic_launch_status();
}
