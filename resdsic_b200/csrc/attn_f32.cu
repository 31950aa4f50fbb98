// Fused shifted-window attention core, fp32 SIMT version.
//
// Replaces (reference layers/win_attention.py): torch.roll(-s) :183,
// window_partition :188, q*scale :94, q@k^T :95, relative-position-bias gather
// :97-101, shift mask build + add (-100) :159-177,103-105, softmax :106,
// attn@v :112, window_reverse :196 and roll(+s) :200 -- none of the
// roll/partition copies are materialised: they are pure addressing here.
//
// One CTA per (window, head); thread i owns query token i: scores in
// registers, K/V of the head staged in shared memory and read as broadcasts.
#include "common.cuh"

namespace {

template <int NTOK>
__global__ void __launch_bounds__(NTOK) win_attn_f32_kernel(const rdsic_attn_desc d) {
  extern __shared__ float smem[];
  const int ws = d.ws, C = d.C, heads = d.heads, dh = C / heads;
  const int ldq = dh + 1;  // odd stride: conflict-free per-thread q rows
  float* qs = smem;                 // [NTOK][dh+1]
  float* ks = qs + NTOK * ldq;      // [NTOK][dh]
  float* vs = ks + NTOK * dh;       // [NTOK][dh]
  int* rid = (int*)(vs + NTOK * dh);  // [NTOK] shift-mask region ids

  const int head = blockIdx.x % heads;
  int win = blockIdx.x / heads;
  const int nWw = d.W / ws, nWh = d.H / ws;
  const int ww = win % nWw;
  win /= nWw;
  const int wh = win % nWh;
  const int b = win / nWh;
  const float scale = d.scale;

  // token -> shifted-frame pixel -> original pixel
  const int i = threadIdx.x;
  const int hy = wh * ws + i / ws, wx = ww * ws + i % ws;
  const int oy = (hy + d.shift) % d.H, ox = (wx + d.shift) % d.W;
  const size_t pix = ((size_t)b * d.H + oy) * d.W + ox;
  {
    int rh = (hy >= d.H - ws) + (hy >= d.H - d.shift);
    int rw = (wx >= d.W - ws) + (wx >= d.W - d.shift);
    rid[i] = d.shift > 0 ? 3 * rh + rw : 0;
  }
  const size_t base = pix * d.qkv.ld + d.qkv.coff + head * dh;
  for (int c = 0; c < dh; ++c) {
    qs[i * ldq + c] = ld_elem(d.qkv.ptr, d.qkv.dtype, base + c) * scale;
    ks[i * dh + c] = ld_elem(d.qkv.ptr, d.qkv.dtype, base + C + c);
    vs[i * dh + c] = ld_elem(d.qkv.ptr, d.qkv.dtype, base + 2 * C + c);
  }
  __syncthreads();

  float s[NTOK];
  const int hi = i / ws, wi = i % ws, myrid = rid[i];
  const int tw = 2 * ws - 1;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NTOK; ++j) {
    float a = 0.f;
    for (int c = 0; c < dh; ++c) a = fmaf(qs[i * ldq + c], ks[j * dh + c], a);
    const int hj = j / ws, wj = j % ws;
    a += d.bias_table[((hi - hj + ws - 1) * tw + (wi - wj + ws - 1)) * heads + head];
    if (rid[j] != myrid) a += -100.0f;
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NTOK; ++j) {
    s[j] = expf(s[j] - mx);
    sum += s[j];
  }
#pragma unroll
  for (int j = 0; j < NTOK; ++j) s[j] = s[j] / sum;
  const size_t obase = pix * d.out.ld + d.out.coff + head * dh;
  for (int c = 0; c < dh; ++c) {
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) o = fmaf(s[j], vs[j * dh + c], o);
    st_elem(d.out.ptr, d.out.dtype, obase + c, o);
  }
}


// ---- v2: specialised on (tokens, head_dim): q / scores / output live in registers, K and V rows are read
// from shared memory as 128-bit broadcasts (4 FMAs per LDS), several heads per CTA, bias table in smem.
template <int NTOK, int DH, int HPB>
__global__ void __launch_bounds__(NTOK * HPB) win_attn_reg_kernel(const rdsic_attn_desc d) {
  extern __shared__ __align__(16) float smem[];
  constexpr int WS = NTOK == 64 ? 8 : 4;
  constexpr int TW = 2 * WS - 1;
  float* ks = smem;                          // [HPB][NTOK][DH]
  float* vs = ks + HPB * NTOK * DH;          // [HPB][NTOK][DH]
  float* tab = vs + HPB * NTOK * DH;         // [TW*TW][HPB]
  int* rid = (int*)(tab + TW * TW * HPB);    // [NTOK]

  const int C = d.C, heads = d.heads;
  const int groups = heads / HPB;
  const int hg = blockIdx.x % groups;
  int win = blockIdx.x / groups;
  const int nWw = d.W / WS, nWh = d.H / WS;
  const int ww = win % nWw;
  win /= nWw;
  const int wh = win % nWh;
  const int b = win / nWh;
  const int hl = threadIdx.x / NTOK, i = threadIdx.x % NTOK;
  const int head = hg * HPB + hl;

  const int hy = wh * WS + i / WS, wx = ww * WS + i % WS;  // shifted-frame pixel of token i
  const int oy = (hy + d.shift) % d.H, ox = (wx + d.shift) % d.W;
  const size_t pix = ((size_t)b * d.H + oy) * d.W + ox;
  if (hl == 0) {
    const int rh = (hy >= d.H - WS) + (hy >= d.H - d.shift), rw = (wx >= d.W - WS) + (wx >= d.W - d.shift);
    rid[i] = d.shift > 0 ? 3 * rh + rw : 0;
  }
  for (int e = threadIdx.x; e < TW * TW * HPB; e += NTOK * HPB)
    tab[e] = d.bias_table[(e / HPB) * heads + hg * HPB + (e % HPB)];

  float q[DH];
  {
    const size_t base = pix * d.qkv.ld + d.qkv.coff + head * DH;
    float* kr = ks + (hl * NTOK + i) * DH;
    float* vr = vs + (hl * NTOK + i) * DH;
    if (d.qkv.dtype == RDSIC_BF16) {
      const __nv_bfloat16* p = (const __nv_bfloat16*)d.qkv.ptr + base;
#pragma unroll
      for (int c = 0; c < DH; c += 8) {
        const uint4 uq = *reinterpret_cast<const uint4*>(p + c);
        const uint4 uk = *reinterpret_cast<const uint4*>(p + C + c);
        const uint4 uv = *reinterpret_cast<const uint4*>(p + 2 * C + c);
        const uint32_t wq[4] = {uq.x, uq.y, uq.z, uq.w}, wk[4] = {uk.x, uk.y, uk.z, uk.w}, wv[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          q[c + 2 * u] = __uint_as_float(wq[u] << 16) * d.scale;
          q[c + 2 * u + 1] = __uint_as_float(wq[u] & 0xFFFF0000u) * d.scale;
          kr[c + 2 * u] = __uint_as_float(wk[u] << 16);
          kr[c + 2 * u + 1] = __uint_as_float(wk[u] & 0xFFFF0000u);
          vr[c + 2 * u] = __uint_as_float(wv[u] << 16);
          vr[c + 2 * u + 1] = __uint_as_float(wv[u] & 0xFFFF0000u);
        }
      }
    } else {
      const float* p = (const float*)d.qkv.ptr + base;
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        const float4 fq = *reinterpret_cast<const float4*>(p + c);
        q[c] = fq.x * d.scale; q[c + 1] = fq.y * d.scale; q[c + 2] = fq.z * d.scale; q[c + 3] = fq.w * d.scale;
        *reinterpret_cast<float4*>(kr + c) = *reinterpret_cast<const float4*>(p + C + c);
        *reinterpret_cast<float4*>(vr + c) = *reinterpret_cast<const float4*>(p + 2 * C + c);
      }
    }
  }
  __syncthreads();

  float s[NTOK];
  const int hi = i / WS, wi = i % WS, myrid = rid[i];
  const float* kh = ks + hl * NTOK * DH;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NTOK; ++j) {
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(kh + j * DH + c);
      a = fmaf(q[c], k4.x, a); a = fmaf(q[c + 1], k4.y, a); a = fmaf(q[c + 2], k4.z, a); a = fmaf(q[c + 3], k4.w, a);
    }
    a += tab[((hi - j / WS + WS - 1) * TW + (wi - j % WS + WS - 1)) * HPB + hl];
    if (rid[j] != myrid) a += -100.0f;
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NTOK; ++j) {
    s[j] = expf(s[j] - mx);
    sum += s[j];
  }
  float o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;
  const float* vh = vs + hl * NTOK * DH;
#pragma unroll
  for (int j = 0; j < NTOK; ++j) {
    const float pj = s[j] / sum;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      const float4 v4 = *reinterpret_cast<const float4*>(vh + j * DH + c);
      o[c] = fmaf(pj, v4.x, o[c]); o[c + 1] = fmaf(pj, v4.y, o[c + 1]);
      o[c + 2] = fmaf(pj, v4.z, o[c + 2]); o[c + 3] = fmaf(pj, v4.w, o[c + 3]);
    }
  }
  const size_t obase = pix * d.out.ld + d.out.coff + head * DH;
  if (d.out.dtype == RDSIC_BF16) {
    __nv_bfloat16* p = (__nv_bfloat16*)d.out.ptr + obase;
#pragma unroll
    for (int c = 0; c < DH; c += 8) {
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(o[c + 2 * u], o[c + 2 * u + 1]);
        w[u] = *reinterpret_cast<uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(p + c) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    float* p = (float*)d.out.ptr + obase;
#pragma unroll
    for (int c = 0; c < DH; c += 4) *reinterpret_cast<float4*>(p + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
  }
}

template <int NTOK, int DH, int HPB>
int launch_reg(const rdsic_attn_desc* d, cudaStream_t stream) {
  constexpr int WS = NTOK == 64 ? 8 : 4;
  const int nblk = d->B * (d->H / WS) * (d->W / WS) * (d->heads / HPB);
  const size_t smem = (size_t)(2 * HPB * NTOK * DH + (2 * WS - 1) * (2 * WS - 1) * HPB) * sizeof(float) + NTOK * sizeof(int);
  auto kern = win_attn_reg_kernel<NTOK, DH, HPB>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<nblk, NTOK * HPB, smem, stream>>>(*d);
  return rdsic_launch_status();
}

}  // namespace

int rdsic_attn_forward_f32(const rdsic_attn_desc* d, cudaStream_t stream) {
  RDSIC_CHECK_ARG(d && d->qkv.ptr && d->out.ptr && d->bias_table);
  RDSIC_CHECK_ARG(d->B > 0 && d->heads > 0 && d->C % d->heads == 0);
  RDSIC_CHECK_ARG(d->ws > 0 && d->H % d->ws == 0 && d->W % d->ws == 0);
  RDSIC_CHECK_ARG(d->shift >= 0 && d->shift < d->ws);  // win_attention.py:144
  RDSIC_CHECK_ARG(!d->qkv.nchw && !d->out.nchw);
  const int ntok = d->ws * d->ws, dh = d->C / d->heads;
  // register-resident specialisations for the reference's two configurations (C=192 w8, C=320 w4; 8 heads)
  const bool vec_ok = d->qkv.ld % 8 == 0 && d->qkv.coff % 8 == 0 && d->out.ld % 8 == 0 && d->out.coff % 8 == 0 &&
                      ((uintptr_t)d->qkv.ptr % 16) == 0 && ((uintptr_t)d->out.ptr % 16) == 0;
  if (vec_ok && ntok == 64 && dh == 24 && d->heads % 2 == 0) return launch_reg<64, 24, 2>(d, stream);
  if (vec_ok && ntok == 16 && dh == 40 && d->heads % 8 == 0) return launch_reg<16, 40, 8>(d, stream);
  const int nblk = d->B * (d->H / d->ws) * (d->W / d->ws) * d->heads;
  const size_t smem = (size_t)ntok * ((dh + 1) + 2 * dh) * sizeof(float) + ntok * sizeof(int);
  if (ntok == 64)
    win_attn_f32_kernel<64><<<nblk, 64, smem, stream>>>(*d);
  else if (ntok == 16)
    win_attn_f32_kernel<16><<<nblk, 16, smem, stream>>>(*d);
  else
    return RDSIC_E_UNSUPPORTED;
  return rdsic_launch_status();
}
