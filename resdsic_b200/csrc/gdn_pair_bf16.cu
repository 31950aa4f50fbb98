// Convolution (or transposed-convolution phase) fused with the GDN / inverse GDN that follows it, on CTA PAIRS
// (tcgen05 cta_group::2) -- the pair form of conv_gdn_tc_kernel<TAIL_GDN | TAIL_IGDN> (reference cnn.py:32-39,46-52,
// layers/gdn.py:62-75):
//
//     x    = conv(in) + bias                         GEMM 1  (implicit GEMM, M = 256 rows over two SMs)
//     norm = beta' + gamma' @ x^2                    GEMM 2  (A = x^2 staged as bf16 in TMEM)
//     out  = x * rsqrt(norm)   |   x * sqrt(norm)    epilogue
//
// Why: every UMMA reads its operands from shared memory at ~64 B/cycle/SM.  A 1-CTA N = 192 tile
// reads 4 KB of A + 6 KB of B per 16-wide K step = 160 cycles against 96 of tensor math: the 1-CTA kernel measures 0.58 -
// 0.65 of the tensor peak and cannot do better.  In a pair each SM reads its own 4 KB of A and HALF of B (3 KB): 110
// cycles.  (B multicast in 1-CTA tiles -- conv_gdn_bf16.cu's MC mode -- halves the L2 bytes, not these reads: no gain.)
//
// Roles per CTA (16 warps): warps 0 / 15 TMA producers (stage parity), warp 1 GEMM-1 issuer and warp 14 GEMM-2 issuer --
// leader CTA only --, warps 2..13 epilogue.  TMEM (C = 192): acc1 [0,C) | P = x^2 bf16 [C, 1.5C) | acc2 [1.5C, 2.5C).
// Per tile: GEMM 1 -> P1 (x kept as packed bf16 in registers, x^2 staged; releases acc1: GEMM 1 of the next tile runs
// under the rest) -> GEMM 2 -> P2.  Barriers: both CTAs' boxes complete on the LEADER's full[s]; the leader's commits are
// multicast to both CTAs' empty[s] / acc1_full / acc2_full; both CTAs' epilogue warps arrive on the leader's p_full
// (P staged, acc1 drained) and acc2_empty (count 24 each).
#include "tc_common.cuh"

namespace {

constexpr int GP_EPI_WARPS = 12;
constexpr int GP_TAIL_WARP = 2 + GP_EPI_WARPS;
constexpr int GP_PRODUCER2_WARP = GP_TAIL_WARP + 1;
constexpr int GP_THREADS = 128 + 32 * GP_EPI_WARPS;
constexpr int GP_PARTS = GP_EPI_WARPS / 4;
constexpr int GP_MAXC = 192;
constexpr int GP_CHUNKS = GP_MAXC / 16 / GP_PARTS;  // 4

struct GdnPairGeom {
  int ns, kb, kiters, kc_last, k2_blocks, kc2_last, b_stage_bytes, g_blk_bytes, p_col, acc2_col;
};

__device__ __forceinline__ void gp_umma_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gp_tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int MODE>  // 1 GDN, 2 inverse GDN
__global__ void __launch_bounds__(GP_THREADS, 1)
gdn_pair_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_g, const rdsic_conv_desc d, const TcGeom g, const GdnPairGeom gg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = A_STAGE_BYTES + gg.b_stage_bytes;
  uint8_t* gamma_s = smem + (size_t)gg.ns * stage_bytes;  // [k2_blocks][C/2 rows x 128 B]: this CTA's half of gamma'
  uint64_t* full_bar = (uint64_t*)(gamma_s + (size_t)gg.k2_blocks * gg.g_blk_bytes);  // [MAX_STAGES] LEADER's copy
  uint64_t* empty_bar = full_bar + MAX_STAGES;   // [MAX_STAGES] local, multicast commit
  uint64_t* acc1_full = empty_bar + MAX_STAGES;  // local, multicast commit
  uint64_t* p_full = acc1_full + 1;              // LEADER's copy: P staged and acc1 drained by both CTAs
  uint64_t* acc2_full = p_full + 1;              // local, multicast commit
  uint64_t* acc2_empty = acc2_full + 1;          // LEADER's copy
  uint64_t* g_full = acc2_empty + 1;             // LEADER's copy: both CTAs' gamma' halves have landed
  uint32_t* tmem_slot = (uint32_t*)(g_full + 1);
  float* bias1_s = (float*)(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);  // [C] conv bias (zeros if none)
  float* bias2_s = bias1_s + GP_MAXC;                                             // [C] beta'

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int C = d.Cout;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_g) : "memory");
    for (int s = 0; s < gg.ns; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(p_full, 2 * GP_EPI_WARPS);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 2 * GP_EPI_WARPS);
    mbar_init(g_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  pdl_wait();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    bias1_s[i] = d.bias ? d.bias[i] : 0.f;
    bias2_s[i] = d.tail_bias[i];
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();
  const uint32_t tmem_base = *tmem_slot;
  const TileWalk wk = make_walk(g);
  const bool leader = wk.crank == 0;

  if (warp == 0 || warp == GP_PRODUCER2_WARP) {
    // ================= TMA producers: ring stages owned by stage parity (even ring depth) =================
    const int pw = warp == 0 ? 0 : 1;
    const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t full0 = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
    const uint32_t empty0 = __shfl_sync(0xffffffffu, smem_u32(empty_bar), 0);
    const uint32_t lead_full0 = mapa_u32(full0, 0u);
    int ns = gg.ns, kb = gg.kb, kiters = gg.kiters, total = g.walk_total, step = wk.step, KW = d.KW, Cin = d.Cin;
    asm volatile("" : "+r"(ns), "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step), "+r"(KW), "+r"(Cin));
    if (pw == 0) {  // gamma' half, once; all bytes complete on the leader's g_full
      const uint32_t gbar = smem_u32(g_full), lead_gbar = mapa_u32(gbar, 0u);
      const uint32_t g0 = __shfl_sync(0xffffffffu, smem_u32(gamma_s), 0);
      if (elect_one()) {
        if (leader) mbar_expect_tx_u32(gbar, 2u * (uint32_t)(gg.k2_blocks * gg.g_blk_bytes));
        for (int k2 = 0; k2 < gg.k2_blocks; ++k2)
          tma_load_2d_2sm_u32(g0 + (uint32_t)(k2 * gg.g_blk_bytes), &tmap_g, lead_gbar, k2 * BK, wk.crank * (C / 2));
      }
      __syncwarp();
    }
    const int kq = kiters / ns, kr = kiters % ns;
    const int n_half = wk.crank * (C / 2);
    int s_base = 0;
    uint32_t ph_base = 0;
    for (int q = wk.first; q < total; q += step) {
      int nt, tx, ty, b;
      tile_of(g, wk, q, nt, tx, ty, b);
      const int x0 = tx * g.TW * d.stride - d.pad_w, y0 = ty * g.TH * d.stride - d.pad_h;
      const int f = pw ^ (s_base & 1);  // this producer's first k-iteration of the tile: the first whose stage has its parity
      const int n_own = (kiters - f + 1) / 2;
      int s = s_base + f;
      uint32_t ph = ph_base;
      if (s >= ns) { s -= ns; ph ^= 1u; }
      int cb = f % kb, tap0 = f / kb;
      int r = tap0 / KW, sx = tap0 % KW;
      int kcol = tap0 * Cin;
      for (int n = 0; n < n_own; ++n) {
        mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
        const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes);
        if (elect_one()) {
          if (leader) mbar_expect_tx_u32(full0 + 8u * (uint32_t)s, 2u * (uint32_t)stage_bytes);
          const uint32_t lbar = lead_full0 + 8u * (uint32_t)s;
          tma_load_4d_2sm_u32(a_dst, &tmap_a, lbar, cb * BK, x0 + sx, y0 + r, b);
          tma_load_2d_2sm_u32(a_dst + A_STAGE_BYTES, &tmap_b, lbar, kcol + cb * BK, n_half);
        }
        __syncwarp();
        s += 2;
        if (s >= ns) { s -= ns; ph ^= 1u; }
        cb += 2;
        while (cb >= kb) {
          cb -= kb;
          kcol += Cin;
          if (++sx == KW) { sx = 0; ++r; }
        }
      }
      ph_base ^= (uint32_t)(kq & 1);
      s_base += kr;
      if (s_base >= ns) { s_base -= ns; ph_base ^= 1u; }
    }
    {  // drain: no multicast commit of the leader may arrive on this CTA's barriers after it has exited
      int s = s_base;
      uint32_t ph = ph_base;
      for (int n = 0; n < ns; ++n) {
        if ((s & 1) == pw) mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
        if (++s == ns) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= GEMM-1 issuer: leader CTA only; K runs sequentially through one accumulator =================
    if (leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t full0 = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
      const uint32_t empty0 = __shfl_sync(0xffffffffu, smem_u32(empty_bar), 0);
      const uint32_t a_u0 = (smem_base & 0x3FFFFu) >> 4, stage_u = (uint32_t)stage_bytes >> 4, b_off = (uint32_t)A_STAGE_BYTES >> 4;
      const uint64_t dconst = make_sw128_desc(0);
      const uint32_t idesc = make_idesc(C, 2 * BM);
      int ns = gg.ns, kb = gg.kb, kiters = gg.kiters, total = g.walk_total, step = wk.step;
      asm volatile("" : "+r"(ns), "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step));
      int s = 0, cb = 0;
      uint32_t ph = 0, lt = 0;
      for (int q = wk.first; q < total; q += step, ++lt) {
        mbar_wait(p_full, (lt & 1u) ^ 1u);  // acc1 drained by BOTH CTAs (phase 1 of the previous tile)
        tcgen05_fence_after();
        for (int n = 0; n < kiters; ++n) {
          mbar_wait_u32(full0 + 8u * (uint32_t)s, ph);
          tcgen05_fence_after();
          const uint64_t da = dconst + (uint64_t)(a_u0 + (uint32_t)s * stage_u), db = da + b_off;
          if (elect_one()) {
            if (cb + 1 != kb || gg.kc_last == 4) {
              umma_bf16_2sm(tbase, da, db, idesc, n > 0 ? 1u : 0u);
              umma_bf16_2sm(tbase, da + 2, db + 2, idesc, 1u);
              umma_bf16_2sm(tbase, da + 4, db + 4, idesc, 1u);
              umma_bf16_2sm(tbase, da + 6, db + 6, idesc, 1u);
            } else {
              for (int k = 0; k < gg.kc_last; ++k) umma_bf16_2sm(tbase, da + 2 * k, db + 2 * k, idesc, (n > 0 || k > 0) ? 1u : 0u);
            }
            tcgen05_commit_2sm_mc_u32(empty0 + 8u * (uint32_t)s, 3);
          }
          __syncwarp();
          if (++s == ns) { s = 0; ph ^= 1u; }
          if (++cb == kb) cb = 0;
        }
        if (elect_one()) tcgen05_commit_2sm_mc_u32(smem_u32(acc1_full), 3);
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == GP_TAIL_WARP) {
    // ================= GEMM-2 issuer: leader CTA only: norm = gamma' @ x^2, A from each SM's own TMEM =================
    if (leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t g_addr = __shfl_sync(0xffffffffu, smem_u32(gamma_s), 0);
      const uint32_t idesc2 = make_idesc(C, 2 * BM);
      int total = g.walk_total, step = wk.step;
      asm volatile("" : "+r"(total), "+r"(step));
      mbar_wait(g_full, 0);
      uint32_t lt = 0;
      for (int q = wk.first; q < total; q += step, ++lt) {
        mbar_wait(p_full, lt & 1u);              // all 24 epilogue warps of the pair have staged x^2
        mbar_wait(acc2_empty, (lt & 1u) ^ 1u);   // ... and drained acc2 (phase 2 of the previous tile)
        tcgen05_fence_after();
        const uint32_t p_t = tbase + (uint32_t)gg.p_col, acc2 = tbase + (uint32_t)gg.acc2_col;
        if (elect_one()) {
          for (int kb2 = 0; kb2 < gg.k2_blocks; ++kb2) {
            const uint64_t dg = make_sw128_desc(g_addr + (uint32_t)(kb2 * gg.g_blk_bytes));
            const int kc2 = kb2 + 1 == gg.k2_blocks ? gg.kc2_last : BK / 16;
            for (int k = 0; k < kc2; ++k)
              gp_umma_ts_2sm(acc2, p_t + (uint32_t)((kb2 * 4 + k) * 8), dg + 2 * k, idesc2, (kb2 | k) ? 1u : 0u);
          }
          tcgen05_commit_2sm_mc_u32(smem_u32(acc2_full), 3);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (both CTAs) =================
    const int q = warp % 4, part = (warp - 2) / 4;
    const int ml = q * 32 + lane;
    const int dy = ml / g.TW, dx = ml % g.TW;
    const int nchunks = C / 16;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t lead_p_full = mapa_u32(smem_u32(p_full), 0u), lead_acc2_empty = mapa_u32(smem_u32(acc2_empty), 0u);
    const uint32_t bias1_a = smem_u32(bias1_s), bias2_a = smem_u32(bias2_s);
    uint32_t lt = 0;
    for (int tq = wk.first; tq < g.walk_total; tq += wk.step, ++lt) {
      int nt, tx, ty, b;
      const bool tile_ok = tile_of(g, wk, tq, nt, tx, ty, b);
      const int oy = ty * g.TH + dy, ox = tx * g.TW + dx;
      const bool row_ok = tile_ok && oy < d.OH && ox < d.OW;
      const size_t pix = ((size_t)b * d.OHt + (oy * d.osy + d.ooy)) * d.OWt + (ox * d.osx + d.oox);
      const uint32_t par = lt & 1u;
      uint32_t xs[GP_CHUNKS][8];  // x as packed bf16, kept for the final multiply

      // ---- phase 1: x = acc1 + bias; keep x, stage x^2
      mbar_wait(acc1_full, par);
      tcgen05_fence_after();
#pragma unroll
      for (int ci = 0; ci < GP_CHUNKS; ++ci) {
        const int j = part + GP_PARTS * ci;
        if (j >= nchunks) break;
        float v[16];
        tmem_ld16(tlane + (uint32_t)(j * 16), v);
        uint32_t st[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 f = lds128(bias1_a + (uint32_t)((j * 16 + 4 * i) * 4));
          const float a0 = v[4 * i] + f.x, a1 = v[4 * i + 1] + f.y, a2 = v[4 * i + 2] + f.z, a3 = v[4 * i + 3] + f.w;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(a0, a1), h1 = __floats2bfloat162_rn(a2, a3);
          xs[ci][2 * i] = *reinterpret_cast<uint32_t*>(&h0);
          xs[ci][2 * i + 1] = *reinterpret_cast<uint32_t*>(&h1);
          __nv_bfloat162 s0 = __floats2bfloat162_rn(a0 * a0, a1 * a1), s1 = __floats2bfloat162_rn(a2 * a2, a3 * a3);
          st[2 * i] = *reinterpret_cast<uint32_t*>(&s0);
          st[2 * i + 1] = *reinterpret_cast<uint32_t*>(&s1);
        }
        gp_tmem_st8(tlane + (uint32_t)gg.p_col + (uint32_t)(j * 8), st);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_u32(lead_p_full);  // x^2 staged, acc1 drained: GEMM 2 / the next GEMM 1 may start

      // ---- phase 2: out = x * (r)sqrt(acc2 + beta')
      mbar_wait(acc2_full, par);
      tcgen05_fence_after();
#pragma unroll
      for (int cp = 0; cp < GP_CHUNKS; cp += 2) {
        const int ja = part + GP_PARTS * cp, jb = ja + GP_PARTS;
        if (ja >= nchunks) break;
        const bool has_b = jb < nchunks;
        uint32_t ua[16], ub[16];
        tmem_ld16_issue(tlane + (uint32_t)gg.acc2_col + (uint32_t)(ja * 16), ua);
        if (has_b) tmem_ld16_issue(tlane + (uint32_t)gg.acc2_col + (uint32_t)(jb * 16), ub);
        tmem_ld_wait();
        tmem_ld_fence(ua);
        if (has_b) tmem_ld_fence(ub);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          if (hb == 1 && !has_b) break;
          const int ci = cp + hb, j = hb ? jb : ja;
          const uint32_t* u = hb ? ub : ua;
          if (!row_ok) continue;
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = lds128(bias2_a + (uint32_t)((j * 16 + 4 * i) * 4));
            v[4 * i] = __uint_as_float(u[4 * i]) + f.x; v[4 * i + 1] = __uint_as_float(u[4 * i + 1]) + f.y;
            v[4 * i + 2] = __uint_as_float(u[4 * i + 2]) + f.z; v[4 * i + 3] = __uint_as_float(u[4 * i + 3]) + f.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = __uint_as_float(xs[ci][i] << 16), x1 = __uint_as_float(xs[ci][i] & 0xFFFF0000u);
            v[2 * i] = MODE == 2 ? x0 * sqrtf(v[2 * i]) : x0 * rsqrtf(v[2 * i]);
            v[2 * i + 1] = MODE == 2 ? x1 * sqrtf(v[2 * i + 1]) : x1 * rsqrtf(v[2 * i + 1]);
          }
          store16(d.out, pix * (size_t)d.out.ld + d.out.coff + j * 16, v, false);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_u32(lead_acc2_empty);  // acc2 drained
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

// Called by rdsic_conv_gdn_forward_bf16 after validation; returns -1 when the layer does not qualify.
int rdsic_gdn_pair_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream) {
  static const int tune = getenv("RDSIC_GDN_PAIR") ? atoi(getenv("RDSIC_GDN_PAIR")) : 1;
  const int C = d->Cout, sms = rdsic_sm_count();
  if (!tune || sms < 2 || (d->tail_mode != 1 && d->tail_mode != 2) || d->tail_n != C || C % BK || C > GP_MAXC || d->groups > 1) return -1;
  const int kb = ceil_div(d->Cin, BK), kiters = d->KH * d->KW * kb;
  if (kiters < 8 || d->KH * d->KW < 2) return -1;  // short-K / pointwise producers are epilogue-bound: the 1-CTA kernel
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return RDSIC_E_UNSUPPORTED;
  TcGeom g = {};
  const int B = d->B, H = d->H, W = d->W, OH = d->OH, OW = d->OW;
  {
    long best = -1;
    for (int tw = 128; tw >= 1; tw /= 2) {
      const int th = BM / tw;
      if (tw * d->stride > 256 || th * d->stride > 256) continue;
      const long area = (long)ceil_div(OW, tw) * tw * ceil_div(OH, th) * th;
      if (best < 0 || area < best) { best = area; g.TW = tw; g.TH = th; }
    }
    if (best < 0) return -1;
  }
  g.tiles_x = ceil_div(OW, g.TW);
  g.tiles_y = ceil_div(OH, g.TH);
  g.BN = C;
  g.n_tiles = 1;
  g.total_tiles = B * g.tiles_y * g.tiles_x;
  g.m_tiles = g.total_tiles;
  if (g.m_tiles < sms) return -1;  // below one wave the pair form has nothing to win
  g.pair = 1;
  g.walk_total = ceil_div(g.m_tiles, 2);
  GdnPairGeom gg;
  gg.kb = kb;
  gg.kiters = kiters;
  gg.kc_last = (d->Cin - (kb - 1) * BK) / 16;
  gg.k2_blocks = C / BK;
  gg.kc2_last = BK / 16;
  gg.b_stage_bytes = (C / 2) * BK * 2;
  gg.g_blk_bytes = (C / 2) * BK * 2;
  gg.p_col = C;
  gg.acc2_col = (C + C / 2 + 31) / 32 * 32;
  if (gg.acc2_col + C > 512 || gg.b_stage_bytes % 1024) return -1;
  const int stage_bytes = A_STAGE_BYTES + gg.b_stage_bytes;
  const size_t fixed = 1024 + (2 * MAX_STAGES + 6) * 8 + 16 + 16 + 2 * GP_MAXC * 4;
  int ns = (int)((227L * 1024 - (long)gg.k2_blocks * gg.g_blk_bytes - (long)fixed) / stage_bytes);
  if (ns > MAX_STAGES) ns = MAX_STAGES;
  ns &= ~1;
  if (ns < 2) return -1;
  gg.ns = ns;
  g.num_stages = ns;

  CUtensorMap ta, tb, tg;
  {
    const cuuint64_t ld_b = (cuuint64_t)d->in.ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld_b, ld_b * W, ld_b * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(g.TW * d->stride), (cuuint32_t)(g.TH * d->stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    void* base = (void*)((const __nv_bfloat16*)d->in.ptr + d->in.coff);
    if (encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return RDSIC_E_ARG;
  }
  auto encode_2d = [&](CUtensorMap* tm, const void* ptr, int K, int rows, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  if (encode_2d(&tb, d->weight, d->KH * d->KW * d->Cin, C, C / 2) != CUDA_SUCCESS) return RDSIC_E_ARG;
  if (encode_2d(&tg, d->tail_weight, C, C, C / 2) != CUDA_SUCCESS) return RDSIC_E_ARG;

  const size_t smem = (size_t)ns * stage_bytes + (size_t)gg.k2_blocks * gg.g_blk_bytes + fixed;
  auto kern = d->tail_mode == 2 ? gdn_pair_tc_kernel<2> : gdn_pair_tc_kernel<1>;
  static bool attr_set[16][3] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool track = dev >= 0 && dev < 16;
  if (!track || !attr_set[dev][d->tail_mode]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    if (track) attr_set[dev][d->tail_mode] = true;
  }
  const int grid = 2 * g.walk_total < (sms & ~1) ? 2 * g.walk_total : (sms & ~1);
  return rdsic_launch(kern, dim3((unsigned)grid), GP_THREADS, smem, stream, true, ta, tb, tg, *d, g, gg);
}
