// Layout / elementwise helpers: NCHW<->NHWC (the reference's permutes at
// win_attention.py:156,204 and the module-boundary layout), dtype casts,
// standalone nn.GELU, and LayerNorm over channels (the stf Swin block,
// reference TCM/tcm.py:214-236).  All memory-bound.
#include "common.cuh"

namespace {

// 32x32 smem-tiled transpose-copy: both the NHWC and the NCHW side are accessed
// with 128-byte coalesced rows whichever direction we go.
__global__ void __launch_bounds__(256) copy_views_kernel(const rdsic_copy_desc d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int HW = d.H * d.W;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const bool src_n = d.src.nchw != 0, dst_n = d.dst.nchw != 0;
  // load: tile[c][p]
  for (int k = ty; k < 32; k += 8) {
    int c, p;
    if (src_n) { c = c0 + k; p = p0 + tx; } else { p = p0 + k; c = c0 + tx; }
    float v = 0.f;
    if (c < d.C && p < HW) {
      size_t idx = src_n ? ((size_t)b * (d.src.ld > 0 ? d.src.ld : d.C) + d.src.coff + c) * HW + p
                         : ((size_t)b * HW + p) * d.src.ld + d.src.coff + c;
      v = ld_elem(d.src.ptr, d.src.dtype, idx);
      if (d.op == 4) {
        const size_t i2 = d.src2.nchw ? ((size_t)b * (d.src2.ld > 0 ? d.src2.ld : d.C) + d.src2.coff + c) * HW + p
                                      : ((size_t)b * HW + p) * d.src2.ld + d.src2.coff + c;
        v = __fadd_rn(v, ld_elem(d.src2.ptr, d.src2.dtype, i2));
      }
      if (d.op == 1) v = gelu_erf(v);
      if (d.op == 2) v = v * v;
      if (d.op == 3) v = fminf(fmaxf(v, 0.f), 1.f);
    }
    if (src_n) tile[k][tx] = v; else tile[tx][k] = v;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    int c, p;
    float v;
    if (dst_n) { c = c0 + k; p = p0 + tx; v = tile[k][tx]; } else { p = p0 + k; c = c0 + tx; v = tile[tx][k]; }
    if (c < d.C && p < HW) {
      size_t idx = dst_n ? ((size_t)b * (d.dst.ld > 0 ? d.dst.ld : d.C) + d.dst.coff + c) * HW + p
                         : ((size_t)b * HW + p) * d.dst.ld + d.dst.coff + c;
      st_elem(d.dst.ptr, d.dst.dtype, idx, v);
    }
  }
}

// LayerNorm over channels (torch semantics: biased variance, eps inside the sqrt), two-pass in registers.
// Sub-warp groups of G lanes own one row (G = 4 for C <= 128, 8 for C <= 256, else 32) so that narrow rows
// (the 48..96-channel stages of stf) do not idle most of a warp; each lane moves 8-element (16/32-byte) vectors.
template <int G>
__global__ void __launch_bounds__(256) layernorm_kernel(const rdsic_ln_desc d) {
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G, gl = threadIdx.x % G;
  const bool active = gid < d.rows;
  const int row = active ? gid : 0;
  const size_t ib = (size_t)row * d.in.ld + d.in.coff, ob = (size_t)row * d.out.ld + d.out.coff;
  constexpr int MAXV = 4;  // up to 4 x 8 channels per lane: C <= 32 * G
  float v[MAXV][8];
  const int nvec = d.C / 8;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = gl + k * G;
    if (vi < nvec) {
      if (d.in.dtype == RDSIC_BF16) {
        const uint4 u = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)d.in.ptr + ib + vi * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[k][2 * i] = __uint_as_float(w[i] << 16);
          v[k][2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>((const float*)d.in.ptr + ib + vi * 8);
        const float4 a1 = *reinterpret_cast<const float4*>((const float*)d.in.ptr + ib + vi * 8 + 4);
        v[k][0] = a0.x; v[k][1] = a0.y; v[k][2] = a0.z; v[k][3] = a0.w;
        v[k][4] = a1.x; v[k][5] = a1.y; v[k][6] = a1.z; v[k][7] = a1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[k][i];
    }
  }
#pragma unroll
  for (int o = G / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / d.C;
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (gl + k * G < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = v[k][i] - mean;
        var = fmaf(t, t, var);
      }
    }
#pragma unroll
  for (int o = G / 2; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / d.C + d.eps);
  if (!active) return;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = gl + k * G;
    if (vi < nvec) {
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = (v[k][i] - mean) * rstd * __ldg(d.gamma + vi * 8 + i) + __ldg(d.beta + vi * 8 + i);
      if (d.out.dtype == RDSIC_BF16) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 h = __floats2bfloat162_rn(o8[2 * i], o8[2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>((__nv_bfloat16*)d.out.ptr + ob + vi * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        *reinterpret_cast<float4*>((float*)d.out.ptr + ob + vi * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>((float*)d.out.ptr + ob + vi * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
    }
  }
}

// generic fallback (any C / alignment): one warp per row, scalar accesses
__global__ void __launch_bounds__(256) layernorm_generic_kernel(const rdsic_ln_desc d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= d.rows) return;
  const size_t ib = (size_t)warp * d.in.ld + d.in.coff, ob = (size_t)warp * d.out.ld + d.out.coff;
  float sum = 0.f;
  for (int c = lane; c < d.C; c += 32) sum += ld_elem(d.in.ptr, d.in.dtype, ib + c);
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / d.C;
  float var = 0.f;
  for (int c = lane; c < d.C; c += 32) {
    float t = ld_elem(d.in.ptr, d.in.dtype, ib + c) - mean;
    var = fmaf(t, t, var);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / d.C + d.eps);
  for (int c = lane; c < d.C; c += 32) {
    float t = (ld_elem(d.in.ptr, d.in.dtype, ib + c) - mean) * rstd;
    st_elem(d.out.ptr, d.out.dtype, ob + c, t * d.gamma[c] + d.beta[c]);
  }
}

// one thread per 8-element (16-byte) chunk of a patch row
__global__ void __launch_bounds__(256) patchify_kernel(const rdsic_patch_desc d) {
  pdl_trigger();
  pdl_wait();
  const int chunks = d.Kp / 8;
  const size_t total = (size_t)d.B * d.OH * d.OW * chunks;
  const int Kreal = d.KH * d.KW * d.C;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % chunks);
    const size_t pix = e / chunks;
    const int ox = (int)(pix % d.OW);
    const size_t t = pix / d.OW;
    const int oy = (int)(t % d.OH), b = (int)(t / d.OH);
    uint32_t w[4];
    float f8[8];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float f[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k = ch * 8 + h * 2 + u;
        float v = 0.f;
        if (k < Kreal) {
          const int tap = k / d.C, c = k - tap * d.C;
          const int r = tap / d.KW, s = tap - r * d.KW;
          const int iy = oy * d.stride - d.pad + r, ix = ox * d.stride - d.pad + s;
          if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) {
            const size_t idx = d.src.nchw ? (((size_t)b * d.C + c) * d.H + iy) * d.W + ix
                                          : (((size_t)b * d.H + iy) * d.W + ix) * d.src.ld + d.src.coff + c;
            v = ld_elem(d.src.ptr, d.src.dtype, idx);
          }
        }
        f[u] = v;
        f8[2 * h + u] = v;
      }
      __nv_bfloat162 hh = __floats2bfloat162_rn(f[0], f[1]);
      w[h] = *reinterpret_cast<uint32_t*>(&hh);
    }
    if (d.dst.dtype == RDSIC_BF16) {
      uint4* dst = reinterpret_cast<uint4*>((__nv_bfloat16*)d.dst.ptr + pix * (size_t)d.dst.ld + d.dst.coff) + ch;
      *dst = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      float4* dst = reinterpret_cast<float4*>((float*)d.dst.ptr + pix * (size_t)d.dst.ld + d.dst.coff) + 2 * ch;
      dst[0] = make_float4(f8[0], f8[1], f8[2], f8[3]);
      dst[1] = make_float4(f8[4], f8[5], f8[6], f8[7]);
    }
  }
}

// Tiled variant for the network's first layer (NCHW fp32 image, few channels): one CTA stages the input
// window of a PT_H x PT_W block of output pixels in shared memory with coalesced row reads, then its threads
// walk the block's (pixel, 16-byte chunk) items in destination order, so every store instruction writes 512
// contiguous bytes.  The gather indices come from a k -> window-offset table built once per CTA.  HBM traffic
// is the image once plus the patch matrix once (the generic kernel above spends ~10x that time on integer
// divisions and scattered 4-byte reads).
constexpr int PT_H = 4, PT_W = 64, PT_MAX_K = 128, PT_MAX_TILE = 9216;  // tile floats (36 KB) + table < 48 KB

__global__ void __launch_bounds__(256) patchify_tiled_kernel(const rdsic_patch_desc d, int tiles_x, int tiles_y) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[PT_MAX_TILE];
  __shared__ int lut[PT_MAX_K];
  const int rows = (PT_H - 1) * d.stride + d.KH, cols = (PT_W - 1) * d.stride + d.KW, pitch = cols | 1;  // odd pitch
  const int plane = rows * pitch;
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y, b = t / tiles_y;
  const int ox0 = tx * PT_W, oy0 = ty * PT_H;
  const int ix0 = ox0 * d.stride - d.pad, iy0 = oy0 * d.stride - d.pad;
  const int Kreal = d.KH * d.KW * d.C;
  for (int k = threadIdx.x; k < d.Kp; k += blockDim.x) {
    int off = -1;
    if (k < Kreal) {
      const int tap = k / d.C, c = k - tap * d.C;
      const int r = tap / d.KW, sx = tap - r * d.KW;
      off = c * plane + r * pitch + sx;
    }
    lut[k] = off;
  }
  const float* src = (const float*)d.src.ptr;
  {
    // (c, ry, rx) advance incrementally: e += blockDim.x is (+dq rows, +dr columns) -- the e % cols / e / cols / q % rows
    // form of the first version spent most of the loop in integer division by run-time values
    const int step = (int)blockDim.x, dq = step / cols, dr = step - dq * cols;
    int rx = (int)threadIdx.x % cols, q = (int)threadIdx.x / cols;
    int ry = q % rows, c = q / rows;
    for (int e = threadIdx.x; e < d.C * rows * cols; e += step) {
      const int iy = iy0 + ry, ix = ix0 + rx;
      float v = 0.f;
      if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) v = __ldg(src + (((size_t)b * d.C + c) * d.H + iy) * d.W + ix);
      tile[c * plane + ry * pitch + rx] = v;
      rx += dr;
      ry += dq;
      if (rx >= cols) { rx -= cols; ++ry; }
      while (ry >= rows) { ry -= rows; ++c; }
    }
  }
  __syncthreads();
  const int chunks = d.Kp / 8;
  __nv_bfloat16* dst = (__nv_bfloat16*)d.dst.ptr;
  const int pstep = (int)blockDim.x / chunks, cstep = (int)blockDim.x - pstep * chunks;
  int ch = (int)threadIdx.x % chunks, p = (int)threadIdx.x / chunks;
  for (int e = threadIdx.x; e < PT_H * PT_W * chunks; e += blockDim.x, p += pstep, ch += cstep) {
    if (ch >= chunks) { ch -= chunks; ++p; }
    const int lx = p % PT_W, ly = p / PT_W;
    const int ox = ox0 + lx, oy = oy0 + ly;
    if (ox >= d.OW || oy >= d.OH) continue;
    const int base = ly * d.stride * pitch + lx * d.stride;
    uint32_t w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int o0 = lut[ch * 8 + 2 * h], o1 = lut[ch * 8 + 2 * h + 1];
      __nv_bfloat162 hh = __floats2bfloat162_rn(o0 >= 0 ? tile[base + o0] : 0.f, o1 >= 0 ? tile[base + o1] : 0.f);
      w[h] = *reinterpret_cast<uint32_t*>(&hh);
    }
    const size_t pix = ((size_t)b * d.OH + oy) * d.OW + ox;
    *(reinterpret_cast<uint4*>(dst + pix * (size_t)d.dst.ld + d.dst.coff) + ch) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// The network's first layer exactly (conv 5x5 s2 p2 on a 3-channel NCHW fp32 image, K = 75 padded to 80): every
// extent a compile-time constant.  ncu of the table-driven tiled kernel above on this layer: 216 us for 113 MB in +
// 377 MB out = 0.30 of the HBM copy bandwidth at 65 % issue-slot use -- instruction-bound (~2100 instructions per
// thread: per-element index arithmetic in the window load, 8 table + 8 data shared-memory loads per 16 output bytes).
// Here one thread owns one output pixel: its 75 inputs are shared-memory loads at immediate offsets from one base,
// packed to 40 registers, transposed through a per-warp shared buffer (176-byte pixel pitch: conflict-free 128-bit
// writes) so that every global store instruction writes 512 contiguous bytes.  ~330 instructions per thread.  The
// window's rows are stored with even and odd columns apart, so the stride-2 reads of a warp (column 2 * lx + s) are
// consecutive words; the transpose buffer takes over the window's shared memory once every thread has its values
// (45 KB per CTA: five CTAs per SM).
constexpr int PF_C = 3, PF_K = 5, PF_S = 2, PF_P = 2, PF_KP = 80, PF_TH = 4, PF_TW = 64;
constexpr int PF_ROWS = (PF_TH - 1) * PF_S + PF_K, PF_COLS = (PF_TW - 1) * PF_S + PF_K, PF_PITCH = 132;
constexpr int PF_PLANE = PF_ROWS * PF_PITCH, PF_OPITCH = 176;  // bytes per pixel in the transpose buffer
constexpr int PF_HALF = 66;  // a window row is stored de-interleaved: even columns at [0, 66), odd ones at [66, 132)
constexpr int PF_TILE_BYTES = PF_C * PF_PLANE * 4, PF_OBUF_BYTES = 8 * 32 * PF_OPITCH;
constexpr int PF_SMEM = PF_OBUF_BYTES > PF_TILE_BYTES ? PF_OBUF_BYTES : PF_TILE_BYTES;  // the transpose buffer re-uses the window's space

__global__ void __launch_bounds__(256) patchify_first_kernel(const rdsic_patch_desc d, int tiles_x, int tiles_y) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t pf_smem[];
  float* tile = reinterpret_cast<float*>(pf_smem);                    // [3][11][132]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* obuf = pf_smem + warp * 32 * PF_OPITCH;  // this warp's 32 pixels x 160 (+16) bytes (after the window is consumed)
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y, b = t / tiles_y;
  const int ox0 = tx * PF_TW, oy0 = ty * PF_TH;
  const int ix0 = ox0 * PF_S - PF_P, iy0 = oy0 * PF_S - PF_P;
  const float* src = (const float*)d.src.ptr + (size_t)b * PF_C * d.H * d.W;
  // ---- input window: one (channel, row) per warp pass, coalesced along x
  for (int row = warp; row < PF_C * PF_ROWS; row += 8) {
    const int c = row / PF_ROWS, ry = row - c * PF_ROWS;
    const int iy = iy0 + ry;
    const bool row_ok = iy >= 0 && iy < d.H;
    const float* sp = src + ((size_t)c * d.H + (row_ok ? iy : 0)) * d.W;
    float* tp = tile + c * PF_PLANE + ry * PF_PITCH;
#pragma unroll
    for (int x = lane; x < PF_COLS; x += 32) {
      const int ix = ix0 + x;
      tp[(x & 1) * PF_HALF + (x >> 1)] = (row_ok && ix >= 0 && ix < d.W) ? __ldg(sp + ix) : 0.f;
    }
  }
  __syncthreads();
  // ---- one output pixel per thread: warp w covers row ly = w / 2, columns (w & 1) * 32 + lane
  const int ly = warp >> 1, lx = ((warp & 1) << 5) + lane;
  const float* base = tile + (ly * PF_S) * PF_PITCH + lx;  // column 2 * lx + s -> (s & 1) * 66 + lx + (s >> 1)
  uint32_t w[PF_KP / 2];
#pragma unroll
  for (int k2 = 0; k2 < PF_KP / 2; ++k2) {
    float f[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = 2 * k2 + u;  // k = (r * 5 + s) * 3 + c: tap-major, channel-minor (the packed weight's K order)
      if (k < PF_K * PF_K * PF_C) {
        const int tap = k / PF_C, c = k - tap * PF_C, r = tap / PF_K, sx = tap - r * PF_K;
        f[u] = base[c * PF_PLANE + r * PF_PITCH + (sx & 1) * PF_HALF + (sx >> 1)];
      } else {
        f[u] = 0.f;
      }
    }
    __nv_bfloat162 hh = __floats2bfloat162_rn(f[0], f[1]);
    w[k2] = *reinterpret_cast<uint32_t*>(&hh);
  }
  __syncthreads();  // every thread holds its pixel: the window's space becomes the transpose buffer
#pragma unroll
  for (int j = 0; j < PF_KP / 8; ++j)
    *reinterpret_cast<uint4*>(obuf + lane * PF_OPITCH + j * 16) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  __syncwarp();
  // ---- the warp's 32 x 160 bytes in destination order: chunk q = pixel * 10 + j
  const int oy = oy0 + ly, oxw = ox0 + ((warp & 1) << 5);
  if (oy < d.OH) {
    __nv_bfloat16* dst = (__nv_bfloat16*)d.dst.ptr + d.dst.coff;
    const size_t pix0 = ((size_t)b * d.OH + oy) * d.OW + oxw;
#pragma unroll
    for (int it = 0; it < PF_KP / 8; ++it) {
      const int q = it * 32 + lane, px = q / (PF_KP / 8), j = q - px * (PF_KP / 8);
      if (oxw + px < d.OW)
        *reinterpret_cast<uint4*>(dst + (pix0 + px) * (size_t)d.dst.ld + j * 8) =
            *reinterpret_cast<const uint4*>(obuf + px * PF_OPITCH + j * 16);
    }
  }
}

}  // namespace

extern "C" int rdsic_patch_forward(const rdsic_patch_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->src.ptr && d->dst.ptr && d->B > 0 && d->C > 0 && d->KH > 0 && d->KW > 0 && d->stride > 0);
  RDSIC_CHECK_ARG(d->Kp % 16 == 0 && d->Kp >= d->KH * d->KW * d->C && !d->dst.nchw);
  if (d->dst.ld % 8 || d->dst.coff % 8 || ((uintptr_t)d->dst.ptr % 16)) return RDSIC_E_ALIGN;
  static const int tune_first = getenv("RDSIC_PATCH_FIRST") ? atoi(getenv("RDSIC_PATCH_FIRST")) : 1;
  if (tune_first && d->src.nchw && d->src.dtype == RDSIC_F32 && d->dst.dtype == RDSIC_BF16 && d->C == PF_C && d->KH == PF_K &&
      d->KW == PF_K && d->stride == PF_S && d->pad == PF_P && d->Kp == PF_KP) {
    static_assert(PF_SMEM <= 48 * 1024, "fits the default dynamic shared memory limit");
    const int tiles_x = ceil_div(d->OW, PF_TW), tiles_y = ceil_div(d->OH, PF_TH);
    return rdsic_launch(patchify_first_kernel, dim3((unsigned)((size_t)d->B * tiles_y * tiles_x)), 256, PF_SMEM, (cudaStream_t)stream,
                        false, *d, tiles_x, tiles_y);
  }
  {
    const int rows = (PT_H - 1) * d->stride + d->KH, cols = (PT_W - 1) * d->stride + d->KW;
    const long tile_floats = (long)d->C * rows * (cols | 1);
    if (d->src.nchw && d->src.dtype == RDSIC_F32 && d->dst.dtype == RDSIC_BF16 && d->Kp <= PT_MAX_K &&
        tile_floats <= PT_MAX_TILE) {
      const int tiles_x = ceil_div(d->OW, PT_W), tiles_y = ceil_div(d->OH, PT_H);
      return rdsic_launch(patchify_tiled_kernel, dim3((unsigned)((size_t)d->B * tiles_y * tiles_x)), 256, 0, (cudaStream_t)stream, false, *d, tiles_x, tiles_y);
    }
  }
  const size_t total = (size_t)d->B * d->OH * d->OW * (d->Kp / 8);
  const size_t want = (total + 255) / 256;
  patchify_kernel<<<(unsigned)(want < (size_t)rdsic_sm_count() * 16 ? want : (size_t)rdsic_sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*d);
  return rdsic_launch_status();
}

extern "C" int rdsic_copy_forward(const rdsic_copy_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->src.ptr && d->dst.ptr && d->B > 0 && d->H > 0 && d->W > 0 && d->C > 0);
  RDSIC_CHECK_ARG(d->op >= 0 && d->op <= 4);
  RDSIC_CHECK_ARG(d->op != 4 || (d->src2.ptr && d->src2.nchw == d->src.nchw));
  RDSIC_CHECK_ARG(d->B <= 65535);
  dim3 grid(ceil_div(d->H * d->W, 32), ceil_div(d->C, 32), d->B);
  return rdsic_launch(copy_views_kernel, grid, 256, 0, (cudaStream_t)stream, false, *d);
}

extern "C" int rdsic_ln_forward(const rdsic_ln_desc* d, rdsic_stream_t stream) {
  RDSIC_CHECK_ARG(d && d->in.ptr && d->out.ptr && d->gamma && d->beta && d->rows > 0 && d->C > 0);
  RDSIC_CHECK_ARG(!d->in.nchw && !d->out.nchw);
  const bool vec = d->C % 8 == 0 && d->in.ld % 8 == 0 && d->in.coff % 8 == 0 && d->out.ld % 8 == 0 && d->out.coff % 8 == 0 &&
                   ((uintptr_t)d->in.ptr % 32) == 0 && ((uintptr_t)d->out.ptr % 32) == 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec && d->C <= 128)
    layernorm_kernel<4><<<ceil_div(d->rows, 64), 256, 0, s>>>(*d);
  else if (vec && d->C <= 256)
    layernorm_kernel<8><<<ceil_div(d->rows, 32), 256, 0, s>>>(*d);
  else if (vec && d->C <= 1024)
    layernorm_kernel<32><<<ceil_div(d->rows, 8), 256, 0, s>>>(*d);
  else
    layernorm_generic_kernel<<<ceil_div(d->rows, 8), 256, 0, s>>>(*d);
  return rdsic_launch_status();
}
