// fp32 implicit-GEMM convolution (SIMT FFMA path) -- the "fp32 mode" of the
// forward pass: symbols / CDF indexes downstream are compared bit-exactly
// against the reference only in this mode (BASELINE.json north_star).
//
// Replaces, per call: nn.Conv2d / ConvTranspose2d phase / nn.Linear / GDN 1x1
// contraction (reference WACNN/utils.py:116-134, layers/layers.py:29-43,
// layers/gdn.py:62-75, layers/win_attention.py:91,113) with the elementwise
// tail fused (RDSIC_EPI_*).  Layout: NHWC activations, weights packed
// [Cout][KH*KW*Cin] tap-major so that a BK=16 slice of K is one contiguous
// 64-byte run of channels of one input pixel.
//
// Tile 128(M) x 64(N) x 16(K), 256 threads, 8x4 accumulators per thread,
// register-prefetched global loads, one smem stage.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

struct RowInfo {
  int b, iy0, ix0;  // input origin of this output position (before adding the tap)
  bool valid;
};

__device__ __forceinline__ RowInfo decode_row(const rdsic_conv_desc& d, int m, int M) {
  RowInfo r;
  r.valid = m < M;
  int mm = r.valid ? m : 0;
  int ox = mm % d.OW;
  int t = mm / d.OW;
  int oy = t % d.OH;
  r.b = t / d.OH;
  r.iy0 = oy * d.stride - d.pad_h;
  r.ix0 = ox * d.stride - d.pad_w;
  return r;
}

template <bool FAST>
__global__ void __launch_bounds__(NT) conv_f32_kernel(const rdsic_conv_desc d) {
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Bs[BK][BS_LD];

  const int M = d.B * d.OH * d.OW;
  const int K = d.KH * d.KW * d.Cin;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const float* __restrict__ in = (const float*)d.in.ptr;
  const float* __restrict__ wt = (const float*)d.weight;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // ---- loader state
  // FAST: thread loads 2 float4 of A (rows ar, ar+64; k offset ak4*4) and 1 float4 of B.
  const int ar = tid / 4, ak4 = tid % 4;
  RowInfo rows[2];
  if (FAST) {
    rows[0] = decode_row(d, m0 + ar, M);
    rows[1] = decode_row(d, m0 + ar + 64, M);
  }
  float4 ra[2], rb;
  float sa[8], sb[4];

  auto load_tile = [&](int k0) {
    if (FAST) {
      const int tap = k0 / d.Cin, c0 = k0 - tap * d.Cin;
      const int r = tap / d.KW, s = tap - r * d.KW;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int iy = rows[i].iy0 + r, ix = rows[i].ix0 + s;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rows[i].valid && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) {
          size_t pix = ((size_t)rows[i].b * d.H + iy) * d.W + ix;
          v = *reinterpret_cast<const float4*>(in + pix * d.in.ld + d.in.coff + c0 + ak4 * 4);
          if (d.a_square) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
        }
        ra[i] = v;
      }
      const int n = n0 + ar;  // ar in [0,64)
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < d.Cout) rb = *reinterpret_cast<const float4*>(wt + (size_t)n * K + k0 + ak4 * 4);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / BK, kk = idx % BK;
        const int k = k0 + kk;
        float v = 0.f;
        RowInfo ri = decode_row(d, m0 + row, M);
        if (ri.valid && k < K) {
          const int tap = k / d.Cin, c = k - tap * d.Cin;
          const int r = tap / d.KW, s = tap - r * d.KW;
          const int iy = ri.iy0 + r, ix = ri.ix0 + s;
          if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) {
            size_t idx_in = d.in.nchw ? (((size_t)ri.b * d.Cin + c) * d.H + iy) * d.W + ix
                                      : (((size_t)ri.b * d.H + iy) * d.W + ix) * d.in.ld + d.in.coff + c;
            v = in[idx_in];
            if (d.a_square) v *= v;
          }
        }
        sa[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * NT;
        const int row = idx / BK, kk = idx % BK;
        const int n = n0 + row, k = k0 + kk;
        sb[i] = (n < d.Cout && k < K) ? wt[(size_t)n * K + k] : 0.f;
      }
    }
  };
  auto store_tile = [&]() {
    if (FAST) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = ar + i * 64;
        As[ak4 * 4 + 0][row] = ra[i].x;
        As[ak4 * 4 + 1][row] = ra[i].y;
        As[ak4 * 4 + 2][row] = ra[i].z;
        As[ak4 * 4 + 3][row] = ra[i].w;
      }
      Bs[ak4 * 4 + 0][ar] = rb.x;
      Bs[ak4 * 4 + 1][ar] = rb.y;
      Bs[ak4 * 4 + 2][ar] = rb.z;
      Bs[ak4 * 4 + 3][ar] = rb.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + i * NT;
        As[idx % BK][idx / BK] = sa[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * NT;
        Bs[idx % BK][idx / BK] = sb[i];
      }
    }
  };

  const int nk = ceil_div(K, BK);
  load_tile(0);
  store_tile();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
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

  // ---- epilogue
  const int HWt = d.OHt * d.OWt;
  const int Cview = d.pixel_shuffle ? d.Cout / 4 : d.Cout;  // channel count of out/res/aux tensors
  const bool need_res = epi_needs_res(d.epilogue);
  const bool need_aux = d.epilogue == RDSIC_EPI_GATE;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const int ox = m % d.OW;
    const int t = m / d.OW;
    const int oy = t % d.OH;
    const int b = t / d.OH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.Cout) continue;
      float v = acc[i][j] + (d.bias ? d.bias[n] : 0.f);
      int py = oy * d.osy + d.ooy, px = ox * d.osx + d.oox, c = n;
      if (d.pixel_shuffle) {
        c = n >> 2;
        py = 2 * oy + ((n >> 1) & 1);
        px = 2 * ox + (n & 1);
      }
      const size_t pix = ((size_t)b * d.OHt + py) * d.OWt + px;
      float res = 0.f, aux = 0.f;
      if (need_res) res = ld_elem(d.res.ptr, d.res.dtype, view_index(d.res, pix, c, HWt, Cview));
      if (need_aux) aux = ld_elem(d.aux.ptr, d.aux.dtype, view_index(d.aux, pix, c, HWt, Cview));
      v = apply_epilogue(d.epilogue, v, res, aux);
      st_elem(d.out.ptr, d.out.dtype, view_index(d.out, pix, c, HWt, Cview), v);
      if (d.out2.ptr) st_elem(d.out2.ptr, d.out2.dtype, view_index(d.out2, pix, c, HWt, Cview), d.out2_square ? v * v : v);
      if (d.out3.ptr) st_elem(d.out3.ptr, d.out3.dtype, view_index(d.out3, pix, c, HWt, Cview), v);
    }
  }
}

}  // namespace

int rdsic_conv_validate(const rdsic_conv_desc* d) {
  RDSIC_CHECK_ARG(d && d->in.ptr && d->weight && d->out.ptr);
  RDSIC_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0);
  RDSIC_CHECK_ARG(d->KH > 0 && d->KW > 0 && d->stride > 0 && d->OH > 0 && d->OW > 0);
  RDSIC_CHECK_ARG(d->OHt > 0 && d->OWt > 0 && d->osy > 0 && d->osx > 0);
  RDSIC_CHECK_ARG(d->pixel_shuffle == 0 || ((d->pixel_shuffle == 2 || d->pixel_shuffle == 3) && d->Cout % 4 == 0));
  RDSIC_CHECK_ARG(d->epilogue >= RDSIC_EPI_NONE && d->epilogue <= RDSIC_EPI_LRP);
  if (epi_needs_res(d->epilogue)) RDSIC_CHECK_ARG(d->res.ptr != nullptr);
  if (d->epilogue == RDSIC_EPI_GATE) RDSIC_CHECK_ARG(d->aux.ptr != nullptr);
  const int groups = d->groups > 1 ? d->groups : 1;
  if (!d->in.nchw) RDSIC_CHECK_ARG(d->in.ld >= d->Cin + d->in.coff + (groups - 1) * d->in_group_stride);
  if (groups > 1) RDSIC_CHECK_ARG(d->Cout % groups == 0 && d->in_group_stride >= 0 && !d->in.nchw && !d->tail_mode && !d->pixel_shuffle);
  return 0;
}

int rdsic_conv_forward_f32(const rdsic_conv_desc* d, cudaStream_t stream) {
  int rc = rdsic_conv_validate(d);
  if (rc) return rc;
  RDSIC_CHECK_ARG(d->in.dtype == RDSIC_F32 && d->w_dtype == RDSIC_F32);
  if (d->groups > 1 || d->pixel_shuffle == 3) return RDSIC_E_UNSUPPORTED;  // grouped / phase-major forms: tensor-core path only
  const int M = d->B * d->OH * d->OW;
  dim3 grid(ceil_div(M, BM), ceil_div(d->Cout, BN));
  const bool fast = !d->in.nchw && d->Cin % BK == 0 && d->in.ld % 4 == 0 && d->in.coff % 4 == 0 &&
                    ((uintptr_t)d->in.ptr % 16 == 0) && ((uintptr_t)d->weight % 16 == 0);
  if (fast)
    conv_f32_kernel<true><<<grid, NT, 0, stream>>>(*d);
  else
    conv_f32_kernel<false><<<grid, NT, 0, stream>>>(*d);
  return rdsic_launch_status();
}
