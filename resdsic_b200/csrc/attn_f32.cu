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

}  // namespace

int rdsic_attn_forward_f32(const rdsic_attn_desc* d, cudaStream_t stream) {
  RDSIC_CHECK_ARG(d && d->qkv.ptr && d->out.ptr && d->bias_table);
  RDSIC_CHECK_ARG(d->B > 0 && d->heads > 0 && d->C % d->heads == 0);
  RDSIC_CHECK_ARG(d->ws > 0 && d->H % d->ws == 0 && d->W % d->ws == 0);
  RDSIC_CHECK_ARG(d->shift >= 0 && d->shift < d->ws);  // win_attention.py:144
  RDSIC_CHECK_ARG(!d->qkv.nchw && !d->out.nchw);
  const int ntok = d->ws * d->ws, dh = d->C / d->heads;
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
