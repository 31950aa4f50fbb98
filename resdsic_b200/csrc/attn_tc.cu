// Fused shifted-window attention on tensor cores (bf16 operands, fp32 accumulate / softmax).
//
// Same contract as attn_f32.cu (reference layers/win_attention.py:94-112,159-200): roll, window partition /
// reverse are addressing; (q*scale) k^T + relative-position bias + shift mask (-100), softmax, @ v.
//
// One CTA = one window x one GROUP of heads (hpc heads per CTA, blockDim = 32 * hpc), one warp = one head.  The
// group's q|k|v columns of the window's tokens are staged once in shared memory with 128-bit coalesced loads (three
// contiguous hpc*DH-channel chunks per token of the qkv GEMM output).  hpc = all heads unless the window would not
// fit in shared memory (see heads_per_cta: splitting for occupancy was measured slower).  Per 16-row
// block of queries:  S = Q K^T with mma.sync m16n8k16 (+ one m16n8k8 for the d mod 16 = 8 tail: d = 24 / 40),
// softmax on the accumulator fragments (quad shuffles), then the probabilities are re-used *in registers*
// as the A operand of P V (the accumulator layout of two n8 tiles is exactly one k16 A fragment);
// V fragments come from ldmatrix.trans on the row-major V rows.  The output overwrites the head's Q columns
// in shared memory and the CTA writes the window back with coalesced 128-bit stores.
//
// (These GEMMs are 64x64x24: far too small for a 128-row tcgen05 tile -- QK^T + PV are 0.6 % of the model's
// FLOPs -- so the warp-level MMA is the right tool here; the 128-row GEMMs all run on tcgen05, conv_bf16.cu.)
#include "common.cuh"

namespace {

__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x1_trans(uint32_t& r0, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ float exp2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- staging helpers shared by the two kernel forms.  ncu (source view) of the first form: 999 of 2903 executed
// instructions per warp were the staging section -- lane-divergent `for (v = lane; v < VPT; v += 32)` loops with their
// reconvergence code around 24 cp.async, and a div / mod bias-table gather -- so the copies are one predicated
// instruction per (token, q|k|v chunk) when a chunk is at most 32 vectors, with per-lane offsets hoisted.
// Issue the cp.async copies of one window: token per warp (tok = warp, warp + nw, ...), 16-byte vector per lane.
__device__ __forceinline__ void attn_stage_window(const __nv_bfloat16* src, const size_t* pixs, __nv_bfloat16* qkv, const int ntok,
                                                  const int LD, const int C, const int VPC, const size_t ld_src, const int warp,
                                                  const int nw, const int lane) {
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(qkv);
  if (VPC <= 32) {
    const bool on = lane < VPC;
    const uint32_t dl = base + 16u * (uint32_t)lane, part_b = 16u * (uint32_t)VPC;
    const __nv_bfloat16* sl = src + lane * 8;
    for (int tok = warp; tok < ntok; tok += nw) {
      const __nv_bfloat16* row = sl + pixs[tok] * ld_src;
      const uint32_t dst = dl + (uint32_t)(tok * LD * 2);
      if (on) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(row) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + part_b), "l"(row + C) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 2u * part_b), "l"(row + 2 * C) : "memory");
      }
    }
  } else {
    for (int tok = warp; tok < ntok; tok += nw) {
      const __nv_bfloat16* row = src + pixs[tok] * ld_src;
      const uint32_t dst0 = base + (uint32_t)(tok * LD * 2);
#pragma unroll
      for (int part = 0; part < 3; ++part)
        for (int v = lane; v < VPC; v += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 16u * (uint32_t)(part * VPC + v)),
                       "l"(row + part * C + v * 8)
                       : "memory");
    }
  }
}
// Write the window's outputs (the Q columns of the staging buffer) back: coalesced 128-bit stores, token per warp.
__device__ __forceinline__ void attn_store_window(__nv_bfloat16* dst, const size_t* pixs, const __nv_bfloat16* qkv, const int ntok,
                                                  const int LD, const int VPT, const size_t ld_dst, const int warp, const int nw,
                                                  const int lane) {
  if (VPT <= 32) {
    const bool on = lane < VPT;
    __nv_bfloat16* dl = dst + lane * 8;
    const __nv_bfloat16* ql = qkv + lane * 8;
    for (int tok = warp; tok < ntok; tok += nw)
      if (on) *reinterpret_cast<uint4*>(dl + pixs[tok] * ld_dst) = *reinterpret_cast<const uint4*>(ql + tok * LD);
  } else {
    for (int tok = warp; tok < ntok; tok += nw) {
      __nv_bfloat16* row = dst + pixs[tok] * ld_dst;
      for (int v = lane; v < VPT; v += 32)
        *reinterpret_cast<uint4*>(row + v * 8) = *reinterpret_cast<const uint4*>(qkv + tok * LD + v * 8);
    }
  }
}
// Relative-position-bias table of the CTA's heads, [head][pos] in shared memory, pre-multiplied by log2(e) (the softmax
// runs in base 2).  Global reads walk the parameter's own [pos][heads] layout (contiguous when the CTA owns all heads).
__device__ __forceinline__ void attn_load_table(float* tab, const float* bias_table, const int heads_total, const int head0,
                                                const int hpc, const int tsz, const int tid, const int nthr) {
  if (hpc == 8) {
    for (int e = tid; e < 8 * tsz; e += nthr) {
      const int pos = e >> 3, h = e & 7;
      tab[h * tsz + pos] = bias_table[pos * heads_total + head0 + h] * 1.4426950408889634f;
    }
  } else {
    for (int e = tid; e < hpc * tsz; e += nthr) {
      const int pos = e / hpc, h = e - pos * hpc;
      tab[h * tsz + pos] = bias_table[pos * heads_total + head0 + h] * 1.4426950408889634f;
    }
  }
}

// One 16-row block of queries of one head (one warp): S = Q K^T, bias / mask / softmax, O = P V; the block's output
// replaces its (already consumed) Q rows in the staging buffer.  `qkv`: [NTOK][LD] bf16, q | k | v chunks of CG channels.
template <int WS, int DH>
__device__ __forceinline__ void attn_row_block(__nv_bfloat16* qkv, const int LD, const int CG, const int head, const int rb,
                                               const float* tb, const int* rid, const bool masked, const float scale,
                                               const int lane) {
  constexpr int NTOK = WS * WS;
  constexpr int NT_S = NTOK / 8;     // n8 tiles of S (keys)
  constexpr int KS_PV = NTOK / 16;   // k16 steps of P V
  constexpr int NT_O = DH / 8;       // n8 tiles of the output
  constexpr int K16 = DH / 16, KTAIL = DH % 16;  // QK^T: K16 k16-steps + (KTAIL == 8) one k8 step
  constexpr int TWD = 2 * WS - 1;
  static_assert(KTAIL == 0 || KTAIL == 8, "head dim must be a multiple of 8");
  const int g = lane / 4, t = lane % 4;
  const __nv_bfloat16* Q = qkv + head * DH;
  const __nv_bfloat16* K = qkv + CG + head * DH;
  const __nv_bfloat16* V = qkv + 2 * CG + head * DH;
  {
    const int i0 = rb * 16 + g, i1 = i0 + 8;
    // ---- S = Q K^T
    float s[NT_S][4];
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < K16; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(Q + i0 * LD + ks * 16 + 2 * t);
      a[1] = *reinterpret_cast<const uint32_t*>(Q + i1 * LD + ks * 16 + 2 * t);
      a[2] = *reinterpret_cast<const uint32_t*>(Q + i0 * LD + ks * 16 + 8 + 2 * t);
      a[3] = *reinterpret_cast<const uint32_t*>(Q + i1 * LD + ks * 16 + 8 + 2 * t);
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
        const __nv_bfloat16* kr = K + (nt * 8 + g) * LD + ks * 16 + 2 * t;
        mma_16816(s[nt], a, *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
      }
    }
    if (KTAIL == 8) {
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(Q + i0 * LD + K16 * 16 + 2 * t);
      const uint32_t a1 = *reinterpret_cast<const uint32_t*>(Q + i1 * LD + K16 * 16 + 2 * t);
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt)
        mma_1688(s[nt], a0, a1, *reinterpret_cast<const uint32_t*>(K + (nt * 8 + g) * LD + K16 * 16 + 2 * t));
    }
    // ---- scale, relative-position bias, shift mask, softmax (rows i0 and i1; a row lives in one quad)
    const int hi0 = i0 / WS, wi0 = i0 % WS, hi1 = i1 / WS, wi1 = i1 % WS;
    const float scale2 = scale * 1.4426950408889634f;  // scores in units of log2(e): exp(v) == exp2(v')
    float m0 = -INFINITY, m1 = -INFINITY;
    // (key j = nt * 8 + 2t + e sits at window row nt * 8 / WS + ..., column (2t + e) % WS: per-thread base pointers, the
    // nt / e parts of the table index are compile-time offsets)
    const float* tb0 = tb + (hi0 + WS - 1) * TWD + (wi0 + WS - 1);
    const float* tb1 = tb + (hi1 + WS - 1) * TWD + (wi1 + WS - 1);
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        const int off = (j / WS) * TWD + (j % WS);
        const float v0 = fmaf(s[nt][e], scale2, tb0[-off]);
        const float v1 = fmaf(s[nt][2 + e], scale2, tb1[-off]);
        s[nt][e] = v0;
        s[nt][2 + e] = v1;
      }
    }
    if (masked) {  // block-uniform: only the windows of the last row / column carry the shift mask (-100)
      const int r0 = rid[i0], r1 = rid[i1];
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int rj = rid[nt * 8 + 2 * t + e];
          if (rj != r0) s[nt][e] += -100.0f * 1.4426950408889634f;
          if (rj != r1) s[nt][2 + e] += -100.0f * 1.4426950408889634f;
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = exp2f_fast(s[nt][e] - m0);
        s[nt][2 + e] = exp2f_fast(s[nt][2 + e] - m1);
        l0 += s[nt][e];
        l1 += s[nt][2 + e];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    // ---- O = P V  (P stays in registers, UNNORMALISED in (0, 1] -- the row's 1 / sum scales the 2 x NT_O outputs
    //      instead of the 2 x NTOK / 4 probabilities; two S n8-tiles form one k16 A fragment)
    float o[NT_O][4];
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KS_PV; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        uint32_t b0, b1;  // B[k = token][n = channel] from row-major V: transposed 8x8 loads
        ldmatrix_x2_trans(b0, b1, V + (kk * 16 + (lane & 15)) * LD + nt * 8);
        mma_16816(o[nt], a, b0, b1);
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      o[nt][0] *= inv0; o[nt][1] *= inv0; o[nt][2] *= inv1; o[nt][3] *= inv1;
    }
    // ---- the block's output replaces its (already consumed) Q rows of this head
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      *reinterpret_cast<uint32_t*>(qkv + i0 * LD + head * DH + nt * 8 + 2 * t) = pack_bf16(o[nt][0], o[nt][1]);
      *reinterpret_cast<uint32_t*>(qkv + i1 * LD + head * DH + nt * 8 + 2 * t) = pack_bf16(o[nt][2], o[nt][3]);
    }
  }
}

// WS: window side (8 or 4); DH: head dim (multiple of 8); one warp per head; `hpc` heads per CTA (blockDim = 32 * hpc,
// gridDim = windows * heads / hpc, the head groups of a window in neighbouring CTAs).
template <int WS, int DH>
__global__ void __launch_bounds__(768) win_attn_tc_kernel(const rdsic_attn_desc d, const int hpc) {
  pdl_trigger();
  pdl_wait();
  constexpr int NTOK = WS * WS;
  const int HEADS = hpc, C = d.heads * DH, CG = HEADS * DH, LD = 3 * CG + 8;  // +8 bf16: conflict-free fragment loads
  const int NTHR = 32 * HEADS;
  const int groups = d.heads / hpc;
  constexpr int RB = NTOK / 16;      // 16-row query blocks
  constexpr int TWD = 2 * WS - 1;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* qkv = (__nv_bfloat16*)smem_raw;              // [NTOK][LD]: q | k | v per token
  size_t* pixs = (size_t*)(qkv + NTOK * LD);                  // [NTOK] pixel index of each token (original frame); 8-byte aligned
  float* tab = (float*)(pixs + NTOK);                         // [HEADS][TWD*TWD]
  int* rid = (int*)(tab + HEADS * TWD * TWD);                 // [NTOK] shift-mask region id

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  int win = blockIdx.x / groups;
  const int head0 = (blockIdx.x % groups) * hpc;  // first head of this CTA's group
  const int nWw = d.W / WS, nWh = d.H / WS;
  const int ww = win % nWw;
  win /= nWw;
  const int wh = win % nWh, b = win / nWh;

  for (int tk = tid; tk < NTOK; tk += NTHR) {
    const int hy = wh * WS + tk / WS, wx = ww * WS + tk % WS;  // shifted-frame position of token tk
    const int oy = (hy + d.shift) % d.H, ox = (wx + d.shift) % d.W;
    pixs[tk] = ((size_t)b * d.H + oy) * d.W + ox;
    const int rh = (hy >= d.H - WS) + (hy >= d.H - d.shift), rw = (wx >= d.W - WS) + (wx >= d.W - d.shift);
    rid[tk] = d.shift > 0 ? 3 * rh + rw : 0;
  }
  __syncthreads();
  attn_stage_window((const __nv_bfloat16*)d.qkv.ptr + head0 * DH + d.qkv.coff, pixs, qkv, NTOK, LD, C, CG / 8, (size_t)d.qkv.ld, warp,
                    HEADS, lane);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // the relative-position-bias table (L2-resident) is fetched while the window's copies are in flight
  attn_load_table(tab, d.bias_table, d.heads, head0, HEADS, TWD * TWD, tid, NTHR);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int head = warp;
  const bool masked = d.shift > 0 && (wh == nWh - 1 || ww == nWw - 1);  // any token of this window in a wrapped region
  const float* tb = tab + head * TWD * TWD;

#pragma unroll
  for (int rb = 0; rb < RB; ++rb) attn_row_block<WS, DH>(qkv, LD, CG, head, rb, tb, rid, masked, d.scale, lane);
  __syncthreads();
  attn_store_window((__nv_bfloat16*)d.out.ptr + head0 * DH + d.out.coff, pixs, qkv, NTOK, LD, CG / 8, (size_t)d.out.ld, warp, HEADS, lane);
}

// Persistent, double-buffered form for the 64-token windows (OFF by default, RDSIC_ATTN_PIPE=1): one CTA per SM walks
// a strided list of windows with TWO staging buffers -- the cp.async copies of window n+1 are in flight while the warps
// compute window n -- and two warps per head (each takes half of the window's row blocks); the bias table is fetched
// once per CTA and a CTA always serves the same head group.  MEASURED (tests/gpu_attn_bench.py, batch 24 x 128 x 192,
// C = 192, w8): 356 us against 298 us of the one-window-per-CTA kernel above (410 against 357 us before the staging
// diet): two independent 8-warp CTAs per SM overlap one window's loads / stores with the other's math just as well, and
// the four block-wide barriers per window of the 16-warp CTA cost more than the per-CTA prologue they save.
template <int WS, int DH>
__global__ void __launch_bounds__(768) win_attn_tc_pipe_kernel(const rdsic_attn_desc d, const int hpc, const int nwin) {
  pdl_trigger();
  pdl_wait();
  constexpr int NTOK = WS * WS, RB = NTOK / 16, RBW = 2, RB_PER = RB / RBW;
  static_assert(RB % RBW == 0, "row blocks split over two warps per head");
  constexpr int TWD = 2 * WS - 1;
  const int HEADS = hpc, C = d.heads * DH, CG = HEADS * DH, LD = 3 * CG + 8;
  const int NW = HEADS * RBW, NTHR = 32 * NW;
  const int groups = d.heads / hpc;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* buf0 = (__nv_bfloat16*)smem_raw;             // [2][NTOK][LD]
  size_t* pixs0 = (size_t*)(buf0 + 2 * NTOK * LD);            // [2][NTOK]
  float* tab = (float*)(pixs0 + 2 * NTOK);                    // [HEADS][TWD*TWD]
  int* rid0 = (int*)(tab + HEADS * TWD * TWD);                // [2][NTOK]
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int head0 = ((int)blockIdx.x % groups) * hpc;
  const int wstride = (int)gridDim.x / groups;
  const int nWw = d.W / WS, nWh = d.H / WS;
  const int VPC = CG / 8;  // 16-byte vectors per token and q / k / v chunk
  const __nv_bfloat16* src = (const __nv_bfloat16*)d.qkv.ptr + head0 * DH + d.qkv.coff;
  __nv_bfloat16* dst = (__nv_bfloat16*)d.out.ptr + head0 * DH + d.out.coff;

  auto prepare = [&](int win, int slot) {  // pixel index / shift-mask region of the window's tokens
    const int ww = win % nWw, wh = (win / nWw) % nWh, b = win / (nWw * nWh);
    for (int tk = tid; tk < NTOK; tk += NTHR) {
      const int hy = wh * WS + tk / WS, wx = ww * WS + tk % WS;
      const int oy = (hy + d.shift) % d.H, ox = (wx + d.shift) % d.W;
      pixs0[slot * NTOK + tk] = ((size_t)b * d.H + oy) * d.W + ox;
      const int rh = (hy >= d.H - WS) + (hy >= d.H - d.shift), rw = (wx >= d.W - WS) + (wx >= d.W - d.shift);
      rid0[slot * NTOK + tk] = d.shift > 0 ? 3 * rh + rw : 0;
    }
  };
  auto issue = [&](int slot) {
    attn_stage_window(src, pixs0 + slot * NTOK, buf0 + slot * NTOK * LD, NTOK, LD, C, VPC, (size_t)d.qkv.ld, warp, NW, lane);
  };

  int win = (int)blockIdx.x / groups;
  if (win < nwin) prepare(win, 0);
  attn_load_table(tab, d.bias_table, d.heads, head0, HEADS, TWD * TWD, tid, NTHR);
  __syncthreads();
  if (win < nwin) issue(0);
  asm volatile("cp.async.commit_group;" ::: "memory");

  const int head = warp % HEADS, rbw = warp / HEADS;
  const float* tb = tab + head * TWD * TWD;
  for (int it = 0; win < nwin; win += wstride, ++it) {
    const int cur = it & 1, nxt = win + wstride;
    if (nxt < nwin) prepare(nxt, cur ^ 1);
    __syncthreads();  // next window's token list visible (its buffer was released by the sync that ended the last round)
    if (nxt < nwin) issue(cur ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty: keeps the group count uniform)
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the newest group: this window has landed
    __syncthreads();
    __nv_bfloat16* qkv = buf0 + cur * NTOK * LD;
    const int ww = win % nWw, wh = (win / nWw) % nWh;
    const bool masked = d.shift > 0 && (wh == nWh - 1 || ww == nWw - 1);
#pragma unroll
    for (int r = 0; r < RB_PER; ++r)
      attn_row_block<WS, DH>(qkv, LD, CG, head, rbw * RB_PER + r, tb, rid0 + cur * NTOK, masked, d.scale, lane);
    __syncthreads();
    attn_store_window(dst, pixs0 + cur * NTOK, qkv, NTOK, LD, VPC, (size_t)d.out.ld, warp, NW, lane);
    __syncthreads();  // buffer / token list `cur` free for the window after next
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// heads per CTA: the whole window in one CTA unless its staging buffer would not fit an SM's shared memory (then the
// heads are split over neighbouring CTAs).  MEASURED (tests/gpu_attn_bench.py, batch 24 x 128 x 192 x C192, w8): 8 heads
// per CTA 385 us, 4 -> 440 us, 2 -> 654 us, 1 -> 1200 us: the kernel is instruction-issue bound (ncu: 69 % of the issue
// slots), and the per-token staging / write-back instructions do not shrink with the group while the CTA count grows --
// more resident CTAs buy nothing.  RDSIC_ATTN_HPC (read once) overrides for such A/B measurements.
int heads_per_cta(int heads, int dh, int ntok) {
  static const int forced = [] {
    const char* e = getenv("RDSIC_ATTN_HPC");
    return e ? atoi(e) : 0;
  }();
  if (forced > 0 && heads % forced == 0 && (forced * dh) % 8 == 0) return forced;
  int hpc = heads;
  while (hpc % 2 == 0 && ((hpc / 2) * dh) % 8 == 0 && (size_t)ntok * (3 * hpc * dh + 8) * 2 > 200 * 1024) hpc /= 2;
  return hpc;
}

template <int WS, int DH>
int launch_tc(const rdsic_attn_desc* d, cudaStream_t stream) {
  constexpr int NTOK = WS * WS, TWD = 2 * WS - 1;
  const int hpc = heads_per_cta(d->heads, DH, NTOK);
  const int LD = 3 * hpc * DH + 8;
  const int nwin = d->B * (d->H / WS) * (d->W / WS), groups = d->heads / hpc;
  static const int pipe_on = getenv("RDSIC_ATTN_PIPE") ? atoi(getenv("RDSIC_ATTN_PIPE")) : 0;
  if constexpr (NTOK >= 64) {
    // persistent double-buffered form (64-token windows): one CTA per SM, two warps per head (measured slower: off)
    const size_t smem2 = (size_t)2 * NTOK * LD * 2 + (size_t)hpc * TWD * TWD * 4 + 2 * NTOK * 4 + 2 * NTOK * 8 + 16;
    const int sms = rdsic_sm_count();
    if (pipe_on && smem2 <= 227 * 1024 && 64 * hpc <= 768 && nwin >= 2 * sms) {
      auto kern = win_attn_tc_pipe_kernel<WS, DH>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) return (int)e;
      int per = sms / groups;
      if (per < 1) per = 1;
      if (per > nwin) per = nwin;
      return rdsic_launch(kern, dim3((unsigned)(per * groups)), 64 * hpc, smem2, stream, false, *d, hpc, nwin);
    }
  }
  const size_t smem = (size_t)NTOK * LD * 2 + (size_t)hpc * TWD * TWD * 4 + NTOK * 4 + NTOK * 8 + 16;
  auto kern = win_attn_tc_kernel<WS, DH>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  return rdsic_launch(kern, dim3((unsigned)(nwin * groups)), 32 * hpc, smem, stream, false, *d, hpc);
}

}  // namespace

// returns RDSIC_E_UNSUPPORTED when the configuration has no tensor-core specialisation (caller then uses
// the register-resident fp32 kernel of attn_f32.cu)
int rdsic_attn_forward_tc(const rdsic_attn_desc* d, cudaStream_t stream) {
  if (d->qkv.dtype != RDSIC_BF16 || d->out.dtype != RDSIC_BF16 || d->heads < 1 || d->heads > 24) return RDSIC_E_UNSUPPORTED;
  if (d->qkv.ld % 8 || d->qkv.coff % 8 || d->out.ld % 8 || d->out.coff % 8 || ((uintptr_t)d->qkv.ptr % 16) ||
      ((uintptr_t)d->out.ptr % 16))
    return RDSIC_E_UNSUPPORTED;
  const int dh = d->C / d->heads;
  if (d->ws == 8 && dh == 24) return launch_tc<8, 24>(d, stream);
  if (d->ws == 4 && dh == 40) return launch_tc<4, 40>(d, stream);
  if (d->ws == 4 && dh == 80) return launch_tc<4, 80>(d, stream);  // cimd "concatenation": the 2M = 640-wide decoder head
  if (d->ws == 4 && dh == 16) return launch_tc<4, 16>(d, stream);  // stf: window 4, head_dim 16, 3..24 heads
  if (d->ws == 4 && dh == 32) return launch_tc<4, 32>(d, stream);
  if (d->ws == 8 && dh == 32) return launch_tc<8, 32>(d, stream);
  return RDSIC_E_UNSUPPORTED;
}
