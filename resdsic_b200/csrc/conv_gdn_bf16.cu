// Convolution (or transposed-convolution phase) FUSED with the GDN / inverse GDN that follows it
// (reference cnn.py:32-39,46-52: `conv -> GDN`, `deconv -> GDN(inverse)`; layers/gdn.py:62-75):
//
//     x    = conv(in) + bias                         GEMM 1 (implicit GEMM, as conv_bf16.cu)
//     norm = beta' + gamma' @ x^2                    GEMM 2 (the GDN 1x1 contraction)
//     out  = x * rsqrt(norm)   |   x * sqrt(norm)    epilogue
//
// in ONE kernel: x never goes to HBM.  The 128 x C accumulator of GEMM 1 sits in TMEM; the epilogue warps
// read it, keep x (bf16) in registers, and write x^2 as packed bf16 straight back into TMEM, from where
// GEMM 2 consumes it as the A operand (tcgen05.mma with A in tensor memory) against gamma', which stays
// resident in shared memory for the whole persistent CTA.  Per image this removes the write of x and x^2 and
// the re-read of both (4 x M x C x 2 bytes) that the unfused pair costs; at H/2 resolution that is 1.2 GB per
// 8 images, more than everything else the two layers move.
//
// TMEM columns (C = 192): [0,C) acc1 | [C, C + C/2) x^2 bf16 (two values per 32-bit column) | [.., +C) acc2.
// Where a second copy of acc1 fits (the ResidualUnit tail, C = 96), GEMM 1 is K-split across two issuer warps
// like conv_bf16.cu (alternate k-iterations, own accumulators [0,C) and [C,2C), summed in phase 1).
#include "tc_common.cuh"

namespace {

constexpr int G_EPI_WARPS = 12;
constexpr int G_ISSUER2_WARP = 2 + G_EPI_WARPS;
constexpr int G_PRODUCER2_WARP = G_ISSUER2_WARP + 1;
constexpr int G_THREADS = 128 + 32 * G_EPI_WARPS;  // two TMA warps, two MMA issuer warps, epilogue warps
// (16 warps = 4 per scheduler: a 17th warp would cap the kernel at 96 registers per thread and spill)
constexpr int G_PARTS = G_EPI_WARPS / 4;
constexpr int MAXC = 192;                    // channel count supported (TMEM: 2.5 C <= 512)
constexpr int MAX_CHUNKS = MAXC / 16 / G_PARTS;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

enum { TAIL_GDN = 1, TAIL_IGDN = 2, TAIL_RU = 3 };

// ---- GEMM 1 main loop of one tile (shared by the single-buffered and the double-buffered flows)
// Halo mode (stride-1 multi-tap layers on maps a 16 x 8 patch tiles exactly): the (16+KH-1) x (8+KW-1) input patch
// of a 64-channel block is loaded ONCE per tile into a small ring of A patches and every tap's A operand is a
// shifted window of it (descriptor start + (r * halo_w + s) rows, SBO = halo row pitch; the hardware applies the
// 128B swizzle on absolute shared-memory address bits, so the windows need no base offset).  The stage ring then
// carries only the weights.  K runs channel-block-major (cb outer, taps inner).  Motivation: ncu of the non-halo
// kernel shows L2->SM traffic 4.6 x the algorithmic bytes at 11.7 TB/s, i.e. at the chip's L2 throughput cap, and
// the nine-fold re-read of A is 40 % of it.  Result: correct, but slower (see rdsic_conv_gdn_forward_bf16).
struct K1Cfg {
  int ns, kb, taps, kiters, kq, kr, kc_last, halo, mc;
  uint32_t full0, empty0, a_u0, stage_u, b_off_u, idesc;
  uint32_t a_full0, a_empty0, patch_u0, patch_u, na_mask, na_shift, pcols;
  int nin, nout, off_in, off_out, unit_out;  // halo walk: inner / outer extents, window offset per inner / outer step
  uint64_t dconst, dconst_halo;
  long long* kts;  // profiling aid: per-k-iteration stamps (loop top, operands ready, issued) of one tile, or null
};
struct K1State {
  int s_base;
  uint32_t ph_base, pa_base;  // ring position / phase and A-patch counter at the start of the tile
};

__device__ __forceinline__ void k1_mmas(uint32_t acc, uint64_t da, uint64_t db, uint32_t idesc, int kc, bool first) {
  if (kc == 4) {
    umma_bf16(acc, da, db, idesc, first ? 0u : 1u);
    umma_bf16(acc, da + 2, db + 2, idesc, 1u);
    umma_bf16(acc, da + 4, db + 4, idesc, 1u);
    umma_bf16(acc, da + 6, db + 6, idesc, 1u);
  } else {
    for (int k = 0; k < kc; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (!first || k > 0) ? 1u : 0u);
  }
}

// whole warp, uniform control flow; `ways` = 1 (single issuer) or 2 (K-split: this issuer owns the stages of its parity)
__device__ __forceinline__ void gemm1_tile(K1State& st, const K1Cfg& c, uint32_t acc1, uint32_t me, int ways) {
  const int f = ways == 2 ? (int)me ^ (st.s_base & 1) : 0;
  int s = st.s_base + f;
  uint32_t ph = st.ph_base;
  if (s >= c.ns) { s -= c.ns; ph ^= 1u; }
  const int n_own = (c.kiters - f + ways - 1) / ways;
  if (!c.halo) {
    int cb = f % c.kb;
    for (int n = 0; n < n_own; ++n) {
      if (c.kts && n < 64) c.kts[4 * n] = clock64();
      mbar_wait_u32(c.full0 + 8u * (uint32_t)s, ph);
      tcgen05_fence_after();
      if (c.kts && n < 64) c.kts[4 * n + 1] = clock64();
      const uint64_t da = c.dconst + (uint64_t)(c.a_u0 + (uint32_t)s * c.stage_u), db = da + c.b_off_u;
      if (elect_one()) {
        k1_mmas(acc1, da, db, c.idesc, cb + 1 != c.kb ? 4 : c.kc_last, n == 0);
        if (c.mc) tcgen05_commit_mc_u32(c.empty0 + 8u * (uint32_t)s, 3);
        else tcgen05_commit_u32(c.empty0 + 8u * (uint32_t)s);
      }
      __syncwarp();
      if (c.kts && n < 64) c.kts[4 * n + 2] = clock64();
      s += ways;
      if (s >= c.ns) { s -= c.ns; ph ^= 1u; }
      cb += ways;
      while (cb >= c.kb) cb -= c.kb;
    }
  } else {
    // k-iteration n = cb * taps + q; the walk over a block's taps has an inner and an outer counter (variant 1:
    // outer r, inner s, one patch per block; variant 2: outer s = patch unit, inner r).  Pure register arithmetic:
    // with the shared-memory port saturated, table reads on this thread's critical path cost 80-500 cycles each.
    int cb = 0, i_in = f, i_out = 0, u_cur = -1;
    while (i_in >= c.nin) { i_in -= c.nin; ++i_out; }
    uint32_t slot_cur = 0, a_cur = 0;
    for (int n = 0; n < n_own; ++n) {
      if (c.kts && n < 64) c.kts[4 * n] = clock64();
      const uint32_t toff = (uint32_t)(i_in * c.off_in + i_out * c.off_out);
      const int u = cb * (int)c.pcols + i_out * c.unit_out;
      if (u != u_cur) {
        if (u_cur >= 0) {  // done with the previous patch (once these MMAs retire)
          if (elect_one()) tcgen05_commit_u32(c.a_empty0 + 8u * slot_cur);
          __syncwarp();
        }
        const uint32_t p = st.pa_base + (uint32_t)u;
        slot_cur = p & c.na_mask;
        mbar_wait_u32(c.a_full0 + 8u * slot_cur, (p >> c.na_shift) & 1u);
        a_cur = c.patch_u0 + slot_cur * c.patch_u;
        u_cur = u;
      }
      if (c.kts && n < 64) c.kts[4 * n + 3] = clock64();  // (A patch ready)
      mbar_wait_u32(c.full0 + 8u * (uint32_t)s, ph);
      tcgen05_fence_after();
      if (c.kts && n < 64) c.kts[4 * n + 1] = clock64();
      const uint64_t da = c.dconst_halo + (uint64_t)(a_cur + toff);
      const uint64_t db = c.dconst + (uint64_t)(c.a_u0 + (uint32_t)s * c.stage_u);
      if (elect_one()) {
        k1_mmas(acc1, da, db, c.idesc, cb + 1 != c.kb ? 4 : c.kc_last, n == 0);
        tcgen05_commit_u32(c.empty0 + 8u * (uint32_t)s);
      }
      __syncwarp();
      if (c.kts && n < 64) c.kts[4 * n + 2] = clock64();
      s += ways;
      if (s >= c.ns) { s -= c.ns; ph ^= 1u; }
      i_in += ways;
      while (i_in >= c.nin) { i_in -= c.nin; ++i_out; }
      if (i_out >= c.nout) { i_out -= c.nout; ++cb; }
    }
    if (elect_one()) tcgen05_commit_u32(c.a_empty0 + 8u * slot_cur);
    __syncwarp();
    st.pa_base += (uint32_t)c.kb * c.pcols;
  }
  st.ph_base ^= (uint32_t)(c.kq & 1);
  st.s_base += c.kr;
  if (st.s_base >= c.ns) { st.s_base -= c.ns; st.ph_base ^= 1u; }
}

struct GdnGeom {
  int w2_bytes;      // resident second-GEMM weight tiles: ceil(K2/64) x [N2 rows x 128 B]
  int p_col, acc2_col;
  int N2, k2_blocks, kc2_last;  // second GEMM: N2 output columns, K2 = N1 in 64-wide blocks
  int ksplit;                   // GEMM 1 split across two issuers / two accumulators
  // dbl (ResidualUnit tail, N2 == 2 C, 4 C + C <= 512): TMEM = two K-split accumulator pairs [b*2C, b*2C + 2C)
  // and two staged operands P[b]; the tail GEMM of tile t writes its accumulator OVER pair t&1 (free once
  // phase 1 has read it).  The two issuer warps run only main loops (K-split as before); the tail GEMM is issued
  // by the first epilogue warp once all twelve have staged their part of P (it would wait for exactly that anyway):
  // the main loop of tile t+1 overlaps phase 1 / tail GEMM / phase 2 of tile t, which the single-buffered flow
  // serialises (measured: 8.4k-cycle main loop + 3.9k exposed per tile, tests/gpu_ru_trace.py).
  int dbl;
  int na;  // halo mode: A patches in the ring (power of two)
};

template <int MODE>
__global__ void __launch_bounds__(G_THREADS, 1)
conv_gdn_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_g, const rdsic_conv_desc d, const TcGeom g,
                   const GdnGeom gg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = g.halo ? g.b_stage_bytes : A_STAGE_BYTES + g.b_stage_bytes;
  uint8_t* gamma_s = smem + (size_t)g.num_stages * stage_bytes;  // 1024-aligned: stage sizes are multiples of 2048
  uint8_t* a_halo = gamma_s + gg.w2_bytes;                       // halo mode: gg.na patches of g.a_halo_bytes
  uint64_t* full_bar = (uint64_t*)(a_halo + (g.halo ? (size_t)gg.na * g.a_halo_bytes : 0));
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc1_full = empty_bar + MAX_STAGES;  // [2] (second entries used in dbl mode)
  uint64_t* acc1_empty = acc1_full + 2;          // [2]
  uint64_t* p_full = acc1_empty + 2;             // [2]
  uint64_t* acc2_full = p_full + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint64_t* g_full = acc2_empty + 1;
  uint64_t* a_full = g_full + 1;    // [4] halo mode: A-patch ring
  uint64_t* a_empty = a_full + 4;   // [4]
  uint32_t* tmem_slot = (uint32_t*)(a_empty + 4);
  uint32_t* tap_off = tmem_slot + 1;  // (48 words reserved)
  // both bias vectors staged once per CTA (16-byte aligned): the epilogue's per-chunk bias loads were its hottest
  // stall (ncu: the first FADD after each bias LDG, stall_long_sb)
  float* bias1_s = (float*)(((uintptr_t)(tap_off + 48) + 15) & ~(uintptr_t)15);  // [C]  conv bias (zeros if none)
  float* bias2_s = bias1_s + MAXC;                                               // [N2] tail bias

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int C = d.Cout;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_g) : "memory");
    for (int s = 0; s < g.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], g.mc ? 2 : 1);  // MC: the issuers of both CTAs of the pair release a stage
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&acc1_full[k], gg.ksplit ? 2 : 1);
      mbar_init(&acc1_empty[k], G_EPI_WARPS);
      mbar_init(&p_full[k], G_EPI_WARPS);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, G_EPI_WARPS);
    mbar_init(g_full, 1);
    for (int k = 0; k < 4; ++k) {
      mbar_init(&a_full[k], 1);
      mbar_init(&a_empty[k], gg.ksplit ? 2 : 1);  // every main-loop issuer walks every patch
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  pdl_wait();  // PDL: barrier init / TMEM allocation above overlap the previous kernel's tail; global memory from here on
  for (int i = threadIdx.x; i < C; i += blockDim.x) bias1_s[i] = d.bias ? d.bias[i] : 0.f;
  for (int i = threadIdx.x; i < gg.N2; i += blockDim.x) bias2_s[i] = d.tail_bias[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (g.mc) cluster_sync_all();  // the peer's barriers are initialised before any multicast copy / commit targets them
  const uint32_t tmem_base = *tmem_slot;
  const TileWalk wk = make_walk(g);

  if (warp == 0 || warp == G_PRODUCER2_WARP) {
    // ================= TMA producers =================
    // whole warp in uniform control flow, one elected lane issues (see elect_one in tc_common.cuh); two
    // producer warps take alternate k-iterations (see conv_bf16.cu)
    {
      const int pw = warp == 0 ? 0 : 1;
      if (pw == 0) {
        if (elect_one()) {
          // second-GEMM weights once per CTA: K2/64 blocks of [N2 rows x 128 B]
          mbar_expect_tx(g_full, (uint32_t)gg.w2_bytes);
          for (int kb = 0; kb < gg.k2_blocks; ++kb) tma_load_2d(gamma_s + (size_t)kb * gg.N2 * 128, &tmap_g, g_full, kb * BK, 0);
        }
        __syncwarp();
      }
      const uint32_t tx_bytes = (g.dbg_skip_load & 1) ? (uint32_t)g.b_stage_bytes : (uint32_t)stage_bytes;
      const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t full0 = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
      const uint32_t empty0 = __shfl_sync(0xffffffffu, smem_u32(empty_bar), 0);
      int ns = g.num_stages, kb = g.kb_per_tap, kiters = d.KH * d.KW * g.kb_per_tap, total = g.walk_total, step = wk.step;
      int KW = d.KW, Cin = d.Cin;
      const uint32_t b_half_off = (uint32_t)(wk.crank * (C / 2) * BK * 2);  // MC: this CTA's half of the B stage
      const int n_half = wk.crank * (C / 2);
      asm volatile("" : "+r"(ns), "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step), "+r"(KW), "+r"(Cin));
      const int kq = kiters / ns, kr = kiters % ns;
      int s_base = 0;
      uint32_t ph_base = 0;
      if (g.halo) {
        // halo mode: the ring carries only weights (k-iteration n = cb * taps + tap); producer 0 also requests
        // the A patch of a channel block when it reaches its first own k-iteration of that block
        const int taps = d.KH * d.KW;
        const uint32_t a_full0 = __shfl_sync(0xffffffffu, smem_u32(a_full), 0);
        const uint32_t a_empty0 = __shfl_sync(0xffffffffu, smem_u32(a_empty), 0);
        const uint32_t patch0 = __shfl_sync(0xffffffffu, smem_u32(a_halo), 0);
        const int pcols = g.halo == 2 ? d.KW : 1, pw_px = g.halo == 2 ? g.TW : g.halo_w;
        const uint32_t a_bytes = (uint32_t)(pw_px * g.halo_h * BK * 2), na_mask = (uint32_t)gg.na - 1u;
        const uint32_t na_shift = gg.na == 4 ? 2u : 1u;
        const int nin = g.halo == 2 ? d.KH : d.KW, nout = g.halo == 2 ? d.KW : d.KH;
        // Patches are requested `ahead` units BEFORE the weight stages of their unit (global unit index U = local tile
        // number * units-per-tile + unit): queued behind eight weight stages in the TMA FIFO they arrived just in
        // time or late (traced: 300-960 cycle waits at every unit change).
        const int upt = kb * pcols;
        const uint32_t ahead = (uint32_t)(gg.na - 1 < 2 ? gg.na - 1 : 2);
        uint32_t pa = 0, lt = 0;  // pa: patches requested so far (producer 0) = next global unit index to request
        const uint32_t n_local = (uint32_t)((total - (int)blockIdx.x + step - 1) / step);  // tiles of this CTA
        for (int tile = blockIdx.x; tile < total; tile += step, ++lt) {  // (halo mode never runs in MC pairs)
          int t = tile;
          const int tx = t % g.tiles_x;
          t /= g.tiles_x;
          const int ty = t % g.tiles_y, b = t / g.tiles_y;
          const int x0 = tx * g.TW - d.pad_w, y0 = ty * g.TH - d.pad_h;
          const int f = pw ^ (s_base & 1);
          const int n_own = (kiters - f + 1) / 2;
          int s = s_base + f;
          uint32_t ph = ph_base;
          if (s >= ns) { s -= ns; ph ^= 1u; }
          int cb = 0, i_in = f, i_out = 0;
          while (i_in >= nin) { i_in -= nin; ++i_out; }
          for (int n = 0; n <= n_own; ++n) {
            // walk position -> (r, s): variant 2 walks s-major (outer = s = patch unit), variant 1 r-major
            const int r = g.halo == 2 ? i_in : i_out, sx = g.halo == 2 ? i_out : i_in;
            const int tidx = r * KW + sx, tunit = g.halo == 2 ? sx : 0;
            if (pw == 0 && n < n_own) {
              // global unit wanted now: this k-iteration's unit + `ahead`, clipped to the CTA's last unit
              uint32_t want = lt * (uint32_t)upt + (uint32_t)(cb * pcols + tunit) + ahead;
              const uint32_t last = n_local * (uint32_t)upt - 1u;
              if (want > last) want = last;
              while (pa <= want) {
                const uint32_t ut = pa / (uint32_t)upt, uu = pa - ut * (uint32_t)upt;  // local tile number, unit in tile
                int t2 = (int)blockIdx.x + (int)ut * step;
                const int tx2 = t2 % g.tiles_x;
                t2 /= g.tiles_x;
                const int ty2 = t2 % g.tiles_y, b2 = t2 / g.tiles_y;
                const int ucb = (int)uu / pcols, ucol = (int)uu - ucb * pcols;
                const uint32_t slot = pa & na_mask;
                mbar_wait_u32(a_empty0 + 8u * slot, ((pa >> na_shift) & 1u) ^ 1u);
                if (elect_one()) {
                  mbar_expect_tx_u32(a_full0 + 8u * slot, a_bytes);
                  tma_load_4d_u32(patch0 + slot * (uint32_t)g.a_halo_bytes, &tmap_a, a_full0 + 8u * slot, ucb * BK,
                                  tx2 * g.TW - d.pad_w + ucol, ty2 * g.TH - d.pad_h, b2);
                }
                __syncwarp();
                ++pa;
              }
            }
            if (n == n_own) break;
            mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            if (elect_one()) {
              mbar_expect_tx_u32(bar, (uint32_t)g.b_stage_bytes);
              tma_load_2d_u32(smem_base + (uint32_t)(s * stage_bytes), &tmap_b, bar, tidx * Cin + cb * BK, 0);
            }
            __syncwarp();
            s += 2;
            if (s >= ns) { s -= ns; ph ^= 1u; }
            i_in += 2;
            while (i_in >= nin) { i_in -= nin; ++i_out; }
            if (i_out >= nout) { i_out -= nout; ++cb; }
          }
          ph_base ^= (uint32_t)(kq & 1);
          s_base += kr;
          if (s_base >= ns) { s_base -= ns; ph_base ^= 1u; }
        }
      } else
      for (int q = wk.first; q < total; q += step) {
        int nt, tx, ty, b;
        tile_of(g, wk, q, nt, tx, ty, b);
        const int x0 = tx * g.TW * d.stride - d.pad_w, y0 = ty * g.TH * d.stride - d.pad_h;
        const int f = pw ^ (s_base & 1);  // ownership by stage parity (even ring depth), see conv_bf16.cu
        const int n_own = (kiters - f + 1) / 2;
        int s = s_base + f;
        uint32_t ph = ph_base;
        if (s >= ns) { s -= ns; ph ^= 1u; }
        int cb = f % kb, tap0 = f / kb;
        int r = tap0 / KW, sx = tap0 % KW;
        int kcol = tap0 * Cin;
        for (int n = 0; n < n_own; ++n) {
          mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
          const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes), bar = full0 + 8u * (uint32_t)s;
          if (elect_one()) {
            mbar_expect_tx_u32(bar, tx_bytes);
            if (!(g.dbg_skip_load & 1)) tma_load_4d_u32(a_dst, &tmap_a, bar, cb * BK, x0 + sx, y0 + r, b);
            if (g.mc) tma_load_2d_mc_u32(a_dst + A_STAGE_BYTES + b_half_off, &tmap_b, bar, kcol + cb * BK, n_half, 3);
            else tma_load_2d_u32(a_dst + A_STAGE_BYTES, &tmap_b, bar, kcol + cb * BK, 0);
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
      if (g.mc) {  // MC drain: no multicast commit of the peer may arrive on this CTA's barriers after it has exited
        int s = s_base;
        uint32_t ph = ph_base;
        for (int n = 0; n < ns; ++n) {
          if ((s & 1) == pw) mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
          if (++s == ns) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == G_ISSUER2_WARP) {
    // ================= MMA issuers (whole warp, elected lane issues) =================
    const uint32_t me = warp == 1 ? 0u : 1u;
    if (!me || gg.ksplit) {
      const uint32_t idesc2 = make_idesc(gg.N2);
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t gamma_addr = __shfl_sync(0xffffffffu, smem_u32(gamma_s), 0);
      K1Cfg c;
      c.ns = g.num_stages; c.kb = g.kb_per_tap; c.taps = d.KH * d.KW; c.kiters = c.taps * c.kb;
      c.kq = c.kiters / c.ns; c.kr = c.kiters % c.ns;
      c.kc_last = (d.Cin - (g.kb_per_tap - 1) * BK) / 16;
      c.halo = g.halo;
      c.mc = g.mc;
      c.full0 = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
      c.empty0 = __shfl_sync(0xffffffffu, smem_u32(empty_bar), 0);
      c.a_u0 = (smem_base & 0x3FFFFu) >> 4;
      c.stage_u = (uint32_t)stage_bytes >> 4;
      c.b_off_u = A_STAGE_BYTES >> 4;
      c.idesc = make_idesc(C);
      c.a_full0 = __shfl_sync(0xffffffffu, smem_u32(a_full), 0);
      c.a_empty0 = __shfl_sync(0xffffffffu, smem_u32(a_empty), 0);
      c.patch_u0 = (__shfl_sync(0xffffffffu, smem_u32(a_halo), 0) & 0x3FFFFu) >> 4;
      c.patch_u = (uint32_t)g.a_halo_bytes >> 4;
      c.na_mask = (uint32_t)gg.na - 1u;
      c.na_shift = gg.na == 4 ? 2u : 1u;
      c.nin = g.halo == 2 ? d.KH : d.KW;
      c.nout = g.halo == 2 ? d.KW : d.KH;
      c.off_in = g.halo == 2 ? g.TW * 8 : 8;            // 16-byte units: one patch row of pixels = 128 B = 8 units
      c.off_out = g.halo == 2 ? 0 : g.halo_w * 8;
      c.unit_out = g.halo == 2 ? 1 : 0;
      c.dconst = make_sw128_desc(0);
      c.dconst_halo = make_sw128_desc_ex(0, (uint32_t)((g.halo == 2 ? g.TW : g.halo_w) * 128), 0);
      c.pcols = g.halo == 2 ? (uint32_t)d.KW : 1u;
      c.kts = nullptr;
      K1State st = {0, 0u, 0u};
      const int ways = gg.ksplit ? 2 : 1;
      int total = g.walk_total, step = wk.step;
      asm volatile("" : "+r"(total), "+r"(step));
      uint32_t lt = 0;
      if (gg.dbl) {
        // ---- dbl mode: both issuers run ONLY main loops (K-split, own accumulator of pair lt & 1); the tail GEMM
        //      is issued by the first epilogue warp
        for (int q = wk.first; q < total; q += step, ++lt) {
          const uint32_t b = lt & 1u, use = (lt >> 1) & 1u;
          long long* tsp = (g.dbg_ts && blockIdx.x == 0 && !me && lane == 0 && lt < 250) ? g.dbg_ts + lt * 16 : nullptr;
          if (tsp) tsp[0] = clock64();
          mbar_wait(&acc1_empty[b], use ^ 1u);  // pair b drained (phase 2 of tile lt - 2 done)
          tcgen05_fence_after();
          if (tsp) tsp[1] = clock64();
          c.kts = (tsp && lt == 3) ? g.dbg_ts + 2048 : nullptr;
          gemm1_tile(st, c, tbase + b * (uint32_t)(2 * C) + me * (uint32_t)C, me, 2);
          if (elect_one()) tcgen05_commit(&acc1_full[b]);
          __syncwarp();
          if (tsp) tsp[2] = tsp[3] = tsp[4] = tsp[5] = clock64();
        }
      } else {
        const uint32_t acc1 = tbase + me * (uint32_t)C, p_t = tbase + (uint32_t)gg.p_col, acc2 = tbase + (uint32_t)gg.acc2_col;
        for (int q = wk.first; q < total; q += step, ++lt) {
          const uint32_t par = lt & 1u;
          // profiling aid (RDSIC_TC_DBG_TS=1, tests/gpu_ru_trace.py): 16 clock64 stamps per tile of CTA 0,
          // slots 0-5 written by issuer 0, slots 8-12 by the first epilogue warp
          long long* tsp = (g.dbg_ts && blockIdx.x == 0 && !me && lane == 0 && lt < 250) ? g.dbg_ts + lt * 16 : nullptr;
          if (tsp) tsp[0] = clock64();
          // ---- GEMM 1: x = conv(in)
          mbar_wait(&acc1_empty[0], par ^ 1u);
          tcgen05_fence_after();
          if (tsp) tsp[1] = clock64();
          c.kts = (tsp && lt == 3) ? g.dbg_ts + 2048 : nullptr;
          gemm1_tile(st, c, acc1, me, ways);
          if (elect_one()) tcgen05_commit(&acc1_full[0]);
          __syncwarp();
          if (me) continue;
          if (tsp) tsp[2] = clock64();
          // ---- GEMM 2: norm = gamma' @ x^2, A operand from TMEM
          if (lt == 0) mbar_wait(g_full, 0);
          mbar_wait(&p_full[0], par);
          if (tsp) tsp[3] = clock64();
          mbar_wait(acc2_empty, par ^ 1u);
          tcgen05_fence_after();
          if (tsp) tsp[4] = clock64();
          if (elect_one()) {
            for (int kb2 = 0; kb2 < gg.k2_blocks; ++kb2) {
              const uint64_t dg = make_sw128_desc(gamma_addr + (uint32_t)(kb2 * gg.N2 * 128));
              const int kc2 = kb2 + 1 == gg.k2_blocks ? gg.kc2_last : BK / 16;
              for (int k = 0; k < kc2; ++k)  // 16 bf16 of K = 8 TMEM columns of the staged operand
                umma_bf16_ts(acc2, p_t + (uint32_t)((kb2 * 4 + k) * 8), dg + 2 * k, idesc2, (kb2 | k) ? 1u : 0u);
            }
            tcgen05_commit(acc2_full);
          }
          __syncwarp();
          if (tsp) tsp[5] = clock64();
        }
      }
    }
    __syncwarp();
  } else if (warp < G_ISSUER2_WARP) {
    // ================= epilogue warps =================
    const int q = warp % 4, part = (warp - 2) / 4;
    const int ml = q * 32 + lane;
    const int dy = ml / g.TW, dx = ml % g.TW;
    const int nchunks1 = C / 16, nchunks2 = gg.N2 / 16;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t lt = 0;
    for (int tq = wk.first; tq < g.walk_total; tq += wk.step, ++lt) {
      int nt, tx, ty, b;
      const bool tile_ok = tile_of(g, wk, tq, nt, tx, ty, b);
      const int oy = ty * g.TH + dy, ox = tx * g.TW + dx;
      const bool row_ok = tile_ok && oy < d.OH && ox < d.OW;
      const size_t pix = ((size_t)b * d.OHt + (oy * d.osy + d.ooy)) * d.OWt + (ox * d.osx + d.oox);
      const uint32_t par = lt & 1u;
      // dbl: accumulator / staged-operand buffer of this tile and the parity of ITS barriers
      const uint32_t bsel = gg.dbl ? (lt & 1u) : 0u, par1 = gg.dbl ? ((lt >> 1) & 1u) : par;
      const uint32_t acc1_c = bsel * (uint32_t)(2 * C), p_c = (uint32_t)gg.p_col + bsel * (uint32_t)(C / 2);
      const uint32_t acc2_c = gg.dbl ? acc1_c : (uint32_t)gg.acc2_col;  // dbl: the tail accumulator overlays pair bsel
      uint32_t xs[MAX_CHUNKS][8];  // GDN: x as packed bf16, kept for the final multiply; RU: prefetched residual

      long long* tsp = (g.dbg_ts && blockIdx.x == 0 && warp == 2 && lane == 0 && lt < 250) ? g.dbg_ts + lt * 16 : nullptr;
      if (tsp) tsp[8] = clock64();
      // ---- phase 1: v = acc1 + bias;  GDN: stage v^2, keep v;  RU: stage gelu(v)
      mbar_wait(&acc1_full[bsel], par1);
      tcgen05_fence_after();
      if (tsp) tsp[9] = clock64();
#pragma unroll
      for (int ci = 0; ci < MAX_CHUNKS; ++ci) {
        const int j = part + G_PARTS * ci;
        if (j >= nchunks1) break;
        float v[16];
        tmem_ld16(tmem_base + lane_off + acc1_c + (uint32_t)(j * 16), v);
        if (gg.ksplit) {  // fixed order: even + odd k-iterations
          float w[16];
          tmem_ld16(tmem_base + lane_off + acc1_c + (uint32_t)(C + j * 16), w);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += w[i];
        }
        {
          const uint32_t bp = smem_u32(bias1_s + j * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = lds128(bp + 16u * i);
            v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
          }
        }
        uint32_t st[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (MODE == TAIL_RU) {
            __nv_bfloat162 h = __floats2bfloat162_rn(gelu_fast(v[2 * i]), gelu_fast(v[2 * i + 1]));
            st[i] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            xs[ci][i] = *reinterpret_cast<uint32_t*>(&h);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i] * v[2 * i], v[2 * i + 1] * v[2 * i + 1]);
            st[i] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
        tmem_st8(tmem_base + lane_off + p_c + (uint32_t)(j * 8), st);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!gg.dbl) mbar_arrive(&acc1_empty[0]);  // acc1 drained: GEMM 1 of the next tile may start
        mbar_arrive(&p_full[bsel]);                // staged operand in place: GEMM 2 may start
      }
      if (tsp) tsp[10] = clock64();
      if (gg.dbl && warp == 2) {
        // tail GEMM: A = staged P[bsel] (TMEM), B = resident W3, D over accumulator pair bsel
        mbar_wait(&p_full[bsel], par1);  // all twelve epilogue warps have staged P and consumed pair bsel
        if (lt == 0) mbar_wait(g_full, 0);
        tcgen05_fence_after();
        const uint32_t idesc2 = make_idesc(gg.N2);
        const uint32_t gamma_addr = smem_u32(gamma_s);
        const uint32_t p_t = tmem_base + p_c, acc2 = tmem_base + acc2_c;
        if (elect_one()) {
          for (int kb2 = 0; kb2 < gg.k2_blocks; ++kb2) {
            const uint64_t dg = make_sw128_desc(gamma_addr + (uint32_t)(kb2 * gg.N2 * 128));
            const int kc2 = kb2 + 1 == gg.k2_blocks ? gg.kc2_last : BK / 16;
            for (int k = 0; k < kc2; ++k)
              umma_bf16_ts(acc2, p_t + (uint32_t)((kb2 * 4 + k) * 8), dg + 2 * k, idesc2, (kb2 | k) ? 1u : 0u);
          }
          tcgen05_commit(acc2_full);
        }
        __syncwarp();
      }
      if (MODE == TAIL_RU) {  // residual x: issue the loads now, they land while GEMM 2 runs
        const __nv_bfloat16* resp = (const __nv_bfloat16*)d.res.ptr + pix * (size_t)d.res.ld + d.res.coff;
#pragma unroll
        for (int ci = 0; ci < MAX_CHUNKS; ++ci) {
          const int j = part + G_PARTS * ci;
          if (j < nchunks2 && row_ok) {
            const Pack8 r = ldg256(resp + j * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) xs[ci][i] = r.w[i];
          }
        }
      }
      // ---- phase 2: GDN: out = x * (r)sqrt(acc2 + beta');  RU: out = gelu(acc2 + bias2 + res)
      mbar_wait(acc2_full, par);
      tcgen05_fence_after();
      if (tsp) tsp[11] = clock64();
      // chunks in pairs: both TMEM loads are in flight before the one wait (the epilogue is latency-bound with
      // three warps per scheduler)
#pragma unroll
      for (int cp = 0; cp < MAX_CHUNKS; cp += 2) {
        const int ja = part + G_PARTS * cp, jb = ja + G_PARTS;
        if (ja >= nchunks2) break;
        const bool has_b = cp + 1 < MAX_CHUNKS && jb < nchunks2;
        uint32_t ua[16], ub[16];
        tmem_ld16_issue(tmem_base + lane_off + acc2_c + (uint32_t)(ja * 16), ua);
        if (has_b) tmem_ld16_issue(tmem_base + lane_off + acc2_c + (uint32_t)(jb * 16), ub);
        tmem_ld_wait();
        tmem_ld_fence(ua);
        if (has_b) tmem_ld_fence(ub);
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        if (hb == 1 && !has_b) break;
        const int ci = cp + hb, j = hb ? jb : ja;
        const uint32_t* u = hb ? ub : ua;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
        if (!row_ok) continue;
        const uint32_t bp = smem_u32(bias2_s + j * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 f = lds128(bp + 16u * i);
          v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x0 = __uint_as_float(xs[ci][i] << 16), x1 = __uint_as_float(xs[ci][i] & 0xFFFF0000u);
          if (MODE == TAIL_RU) {
            v[2 * i] = gelu_fast(v[2 * i] + x0);
            v[2 * i + 1] = gelu_fast(v[2 * i + 1] + x1);
          } else {
            v[2 * i] = MODE == TAIL_IGDN ? x0 * sqrtf(v[2 * i]) : x0 * rsqrtf(v[2 * i]);
            v[2 * i + 1] = MODE == TAIL_IGDN ? x1 * sqrtf(v[2 * i + 1]) : x1 * rsqrtf(v[2 * i + 1]);
          }
        }
        store16(d.out, pix * (size_t)d.out.ld + d.out.coff + j * 16, v, false);
      }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(gg.dbl ? &acc1_empty[bsel] : acc2_empty);  // dbl: pair bsel (acc1 + overlaid acc2) is free
      if (tsp) tsp[12] = clock64();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

int rdsic_conv_validate(const rdsic_conv_desc* d);
int rdsic_ru_pair_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream);   // ru_pair_bf16.cu
int rdsic_gdn_pair_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream);  // gdn_pair_bf16.cu

// conv (+bias) followed by a second, pointwise GEMM in one launch (d->tail_mode):
//   1 GDN / 2 inverse GDN: tail_weight = packed bf16 gamma' [C][C], tail_bias = fp32 beta' [C], tail_n = C;
//   3 ResidualUnit tail (layers/layers.py:58-71): out = gelu(W3 @ gelu(conv + bias) + b3 + res),
//     tail_weight = packed bf16 [tail_n][C], tail_bias = b3, res = the unit's input.
int rdsic_conv_gdn_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream) {
  const int C = d->Cout;
  const int N2 = d->tail_n;
  rdsic_conv_desc chk = *d;
  chk.Cout = N2 > C ? N2 : C;  // validation of the output views is against the final channel count
  int rc = rdsic_conv_validate(&chk);
  if (rc) return rc;
  RDSIC_CHECK_ARG(d->tail_weight && d->tail_bias && d->tail_mode >= TAIL_GDN && d->tail_mode <= TAIL_RU);
  RDSIC_CHECK_ARG(d->epilogue == RDSIC_EPI_NONE && !d->aux.ptr && !d->out2.ptr && !d->out3.ptr);
  RDSIC_CHECK_ARG(C % 16 == 0 && N2 % 16 == 0 && N2 >= 32 && N2 <= MAXC && C + C / 2 + N2 <= 512);
  RDSIC_CHECK_ARG(d->Cin % 16 == 0 && !d->pixel_shuffle && !d->a_square && d->groups <= 1);
  if (d->tail_mode == TAIL_RU) {
    RDSIC_CHECK_ARG(d->res.ptr && d->res.dtype == RDSIC_BF16 && !d->res.nchw);
    if (d->res.ld % 16 || d->res.coff % 16 || ((uintptr_t)d->res.ptr % 32)) return RDSIC_E_ALIGN;
  } else {
    RDSIC_CHECK_ARG(!d->res.ptr && N2 == C && C % BK == 0);
  }
  RDSIC_CHECK_ARG(d->in.dtype == RDSIC_BF16 && !d->in.nchw && !d->out.nchw && (d->stride == 1 || d->stride == 2));
  if (d->in.ld % 8 || d->in.coff % 8 || ((uintptr_t)d->in.ptr % 16) || ((uintptr_t)d->weight % 16) ||
      ((uintptr_t)d->tail_weight % 16) || ((uintptr_t)d->tail_bias % 16) || (d->bias && ((uintptr_t)d->bias % 16)) ||
      d->out.ld % 16 || d->out.coff % 16 || ((uintptr_t)d->out.ptr % 32))
    return RDSIC_E_ALIGN;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return RDSIC_E_UNSUPPORTED;
  if (d->tail_mode == TAIL_RU) {  // ResidualUnit tail on CTA pairs with resident weights where the layer qualifies
    rc = rdsic_ru_pair_forward_bf16(d, stream);
    if (rc != -1) return rc;
  } else {  // long-K conv / deconv phase + GDN on CTA pairs (half the B reads per SM)
    rc = rdsic_gdn_pair_forward_bf16(d, stream);
    if (rc != -1) return rc;
  }

  TcGeom g = {};
  int B = d->B, H = d->H, W = d->W, OH = d->OH, OW = d->OW;
  const bool flat = d->KH == 1 && d->KW == 1 && d->stride == 1 && d->pad_h == 0 && d->pad_w == 0 && d->osy == 1 &&
                    d->osx == 1 && d->ooy == 0 && d->oox == 0 && OH == H && OW == W && d->OHt == OH && d->OWt == OW;
  rdsic_conv_desc dd = *d;
  if (flat) {
    W = OW = B * H * W;
    H = OH = 1;
    B = 1;
    dd.B = 1; dd.H = 1; dd.W = W; dd.OH = 1; dd.OW = OW; dd.OHt = 1; dd.OWt = OW;
    g.TH = 1; g.TW = 128;
  } else {
    long best = -1;
    for (int tw = 128; tw >= 1; tw /= 2) {
      const int th = BM / tw;
      const long area = (long)ceil_div(OW, tw) * tw * ceil_div(OH, th) * th;
      if (best < 0 || area < best) { best = area; g.TW = tw; g.TH = th; }
    }
  }
  // halo mode (see gemm1_tile): stride-1 multi-tap layers on maps that a 16 x 8 patch tiles exactly.  OFF by default:
  // correct (tests/test_gpu_ops.py::test_kernel_mode_switches_in_subprocess) but measured SLOWER (ResidualUnit main loop
  // 670 vs 520 cycles per k-iteration): it removes L2 traffic, but the bound that remains is the shared-memory
  // port -- the MMAs read A and B from shared memory every k-step whatever wrote them, and windows that do not
  // start on a 1024-byte swizzle atom were the suspect -- but variant 2 below (atom-aligned windows) is just as slow
  // (12.7k vs 10.3k cycles per tile), so misalignment is not the cause.  See DESIGN.md section 4.
  //   RDSIC_GDN_HALO=1: one (TH+KH-1) x (TW+KW-1) patch per channel block, windows at arbitrary row offsets;
  //   RDSIC_GDN_HALO=2: KW column-shifted copies of a (TH+KH-1) x TW patch per channel block, so that every tap's
  //                     window starts on a 1024-byte swizzle atom (offset r * TW rows) with the standard SBO.
  static const int tune_halo = getenv("RDSIC_GDN_HALO") ? atoi(getenv("RDSIC_GDN_HALO")) : 0;
  if (tune_halo && !flat && d->stride == 1 && d->KH * d->KW >= 2 && d->KH <= 3 && d->KW <= 3 && OH % 16 == 0 && OW % 8 == 0 &&
      (tune_halo != 2 || d->KH >= 2)) {
    g.halo = tune_halo == 2 ? 2 : 1;
    g.TH = 16;
    g.TW = 8;
    g.halo_w = g.TW + d->KW - 1;
    g.halo_h = g.TH + d->KH - 1;
    g.a_halo_bytes = (BK * 2 * (g.halo == 2 ? g.TW : g.halo_w) * g.halo_h + 1023) / 1024 * 1024;
  }
  g.tiles_x = ceil_div(OW, g.TW);
  g.tiles_y = ceil_div(OH, g.TH);
  g.BN = C;
  g.n_tiles = 1;
  g.total_tiles = B * g.tiles_y * g.tiles_x;
  g.m_tiles = g.total_tiles;
  static const int tune_mc = getenv("RDSIC_TC_MC") ? atoi(getenv("RDSIC_TC_MC")) : 0;  // measured: no gain, off
  int sms = rdsic_sm_count();
  g.mc = tune_mc && !g.halo && g.m_tiles >= 2 && sms >= 2 && (tune_mc != 2 || g.total_tiles >= sms);  // see TcGeom
  g.walk_total = g.mc ? ceil_div(g.m_tiles, 2) : g.total_tiles;
  g.kb_per_tap = ceil_div(d->Cin, BK);
  g.num_k_iters = d->KH * d->KW * g.kb_per_tap;
  g.b_stage_bytes = C * BK * 2;
  g.tmem_cols = 512;
#ifdef RDSIC_DEBUG
  static const int dbg_skip = getenv("RDSIC_TC_DBG_SKIPG") ? atoi(getenv("RDSIC_TC_DBG_SKIPG")) : 0;
  g.dbg_skip_load = dbg_skip;
  {
    extern long long* g_dbg_ts;
    extern int g_dbg_host;
    static long long* ts_buf = nullptr;
    if (getenv("RDSIC_TC_DBG_TS") && !g_dbg_host && !ts_buf) {
      if (!g_dbg_ts) {
        cudaMalloc(&g_dbg_ts, 4096 * sizeof(long long));
        cudaMemset(g_dbg_ts, 0, 4096 * sizeof(long long));
      }
      ts_buf = g_dbg_ts;
    }
    g.dbg_ts = ts_buf;
  }
#endif
  GdnGeom gg;
  gg.N2 = N2;
  gg.k2_blocks = ceil_div(C, BK);
  gg.kc2_last = (C - (gg.k2_blocks - 1) * BK) / 16;
  gg.w2_bytes = gg.k2_blocks * N2 * 128;
  static const int tune_ksplit = getenv("RDSIC_TC_KSPLIT") ? atoi(getenv("RDSIC_TC_KSPLIT")) : 1;
  gg.ksplit = tune_ksplit && g.num_k_iters >= 4 && (2 * C + C / 2 + 31) / 32 * 32 + N2 <= 512;
  static const int tune_dbl = getenv("RDSIC_RU_DBL") ? atoi(getenv("RDSIC_RU_DBL")) : 1;
  gg.dbl = tune_dbl && gg.ksplit && d->tail_mode == TAIL_RU && N2 == 2 * C && 5 * C <= 512;
  gg.p_col = gg.dbl ? 4 * C : C * (1 + gg.ksplit);
  gg.acc2_col = gg.dbl ? 0 : (gg.p_col + C / 2 + 31) / 32 * 32;
  RDSIC_CHECK_ARG(gg.acc2_col + N2 <= 512);
  gg.na = 0;
  int stage_bytes = A_STAGE_BYTES + g.b_stage_bytes;
  static const int tune_smem_kb = getenv("RDSIC_GDN_SMEM_KB") ? atoi(getenv("RDSIC_GDN_SMEM_KB")) : 200;
  int stages = (tune_smem_kb * 1024 - gg.w2_bytes) / stage_bytes;
  if (g.halo) {
    stage_bytes = g.b_stage_bytes;
    gg.na = C <= 96 ? 4 : 2;
    stages = (220 * 1024 - gg.w2_bytes - gg.na * g.a_halo_bytes) / stage_bytes;
  }
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  stages &= ~1;  // even ring depth: stage parity = owner (two producers / two issuers)
  if (stages < 2) {
    if (!g.halo) return RDSIC_E_ARG;
    return RDSIC_E_UNSUPPORTED;  // (not reachable for the shapes the models use)
  }
  g.num_stages = stages;
  RDSIC_CHECK_ARG(g.TW * d->stride <= 256 && g.TH * d->stride <= 256);

  CUtensorMap ta, tb, tg;
  {
    const cuuint64_t ld_b = (cuuint64_t)d->in.ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld_b, ld_b * W, ld_b * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(g.TW * d->stride), (cuuint32_t)(g.TH * d->stride), 1};
    if (g.halo) { box[1] = (cuuint32_t)(g.halo == 2 ? g.TW : g.halo_w); box[2] = (cuuint32_t)g.halo_h; }
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
  if (encode_2d(&tb, d->weight, d->KH * d->KW * d->Cin, C, g.mc ? C / 2 : C) != CUDA_SUCCESS) return RDSIC_E_ARG;  // MC: half per CTA
  if (encode_2d(&tg, d->tail_weight, C, N2, N2) != CUDA_SUCCESS) return RDSIC_E_ARG;

#ifdef RDSIC_DEBUG
  {  // debug: barrier-timeout log (see tc_common.cuh)
    extern long long* g_dbg_ts;
    extern int g_dbg_host;
    static bool log_set = false;
    if (g_dbg_host && !log_set) {
      tc_set_timeout_log(g_dbg_ts);
      log_set = true;
    }
  }
#endif
  const size_t smem = (size_t)stages * stage_bytes + gg.w2_bytes + (size_t)gg.na * g.a_halo_bytes + 1024 +
                      (2 * MAX_STAGES + 20) * 8 + 16 + 192 + 16 + 2 * MAXC * 4;
  auto kern = d->tail_mode == TAIL_RU ? conv_gdn_tc_kernel<TAIL_RU>
              : d->tail_mode == TAIL_IGDN ? conv_gdn_tc_kernel<TAIL_IGDN> : conv_gdn_tc_kernel<TAIL_GDN>;
  static bool attr_set[16][4] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool track = dev >= 0 && dev < 16;
  if (!track || !attr_set[dev][d->tail_mode]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    if (track) attr_set[dev][d->tail_mode] = true;
  }
  if (g.mc) {
    const int grid = 2 * g.walk_total < (sms & ~1) ? 2 * g.walk_total : (sms & ~1);
    return rdsic_launch(kern, dim3((unsigned)grid), G_THREADS, smem, stream, true, ta, tb, tg, dd, g, gg);
  }
  const int grid = g.total_tiles < sms ? g.total_tiles : sms;
  return rdsic_launch(kern, dim3((unsigned)grid), G_THREADS, smem, stream, false, ta, tb, tg, dd, g, gg);
}
