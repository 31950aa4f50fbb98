// Fused ResidualUnit tail (reference layers/layers.py:58-71: conv3x3 -> GELU -> conv1x1 -> + x -> GELU) on CTA PAIRS
// (tcgen05 cta_group::2) with EVERY weight resident in shared memory and the 3x3's input read ONCE:
//
//     t   = gelu(conv3x3(in) + b2)                       GEMM 1  (implicit GEMM, M = 256 rows over two SMs)
//     out = gelu(W3 @ t + b3 + x)                        GEMM 2  (A = t staged as bf16 in TMEM)
//
// Why a second form of conv_gdn_tc_kernel<TAIL_RU>: ncu of the 1-CTA kernel (profiles/r2_ncu_ru_*.csv) shows its main
// loop at the chip-wide L2->SM cap -- 2.6 GB per launch at 11.1 TB/s, 567 KB per 128-row tile: every tile re-streams the
// 3x3 weights (216 KB) and reads its input nine times, once per filter tap (216 KB).  Here
//   * each SM of a pair holds only HALF of the weight rows (N/2 = 48 of 96; tail: 96 of 192), so the whole weight set
//     (18 k-blocks x 6 KB + 2 x 12 KB = 132 KB) is resident, loaded once per persistent CTA;
//   * the input of a tile is ONE (16+2) x (8+2)-pixel halo patch per 64-channel block (23 KB); the A operand of tap
//     (r, s) is a shifted window of it (descriptor start + (r * 10 + s) pixel rows, SBO = the patch's row pitch; the
//     hardware applies the 128B swizzle on absolute shared-memory address bits, so the windows need no base offset);
//   * with both operands resident there is no per-k-iteration barrier at all: an issuer waits once per tile for its
//     channel block's patch and then issues its 9 taps back to back.
// Per 128-row tile L2->SM carries 46 KB of patches + 49 KB of residual instead of 567 KB.
//
// Roles per CTA (16 warps): warp 0 TMA producer; warp 1 MMA issuer and warp 14 tail-GEMM issuer -- active in the LEADER
// CTA only, issuing cta_group::2 UMMAs (256 x 96 x 16) for both SMs; warps 2..13 epilogue; warp 15 idle.  With both
// operands resident one issuer is enough (54 MMAs per tile, ~5k cycles, hidden behind the epilogue), so K runs through
// ONE accumulator (channel block 0's nine taps, then channel block 1's).
// TMEM (512 columns per SM, C = 96): acc1[b] at b*C (two buffers), staged operands P[b] at 2C + b*C/2, the tail
// accumulator acc2 (2C columns, one buffer) at 3C: 480 columns.  Measured: 155 us per launch (1-CTA kernel: 232).
// Epilogue order (skewed by one tile): P1(0); then per tile lt: P1(lt+1), P2(lt).  P1 stages gelu(acc1 + b2) as the tail
// operand, P2 finishes out = gelu(acc2 + b3 + x).  The tail GEMM of lt (remote arrivals of 24 warps, 6 MMAs behind
// whatever main loop is queued, the multicast commit: 3-4k cycles, exposed in the unskewed order) runs under P1(lt+1);
// the main loop of lt+2 is released by P1(lt) (acc1[b] drained) and runs under P2(lt-1) .. P2(lt).
// Barriers: both CTAs' patches complete on the LEADER's a_full[slot]; the leader's commits are multicast to both CTAs'
// a_empty[slot] / acc1_full[b] / acc2_full; both CTAs' epilogue warps arrive on the leader's p_full[b] (P[b] staged and
// acc1[b] drained: releases the tail GEMM of lt and the main loop of lt+2) and acc2_empty (count 24 each).
// Patch slots: 2 (one per channel block: the main loop runs ~10k cycles ahead of the epilogue, it does not need more),
// filled and consumed in one order by one producer / issuer.  The shared memory they free holds the RESIDUAL tile
// (128 rows x 192 channels, three swizzled 16 KB blocks loaded by a fourth TMA-issuing warp): per-lane LDG.256 of the
// residual -- 32 sectors per instruction, issued in bursts between compute phases -- cost 39 of 230 us per launch.
// (An earlier form with a ring shared by two issuers let one wait on a barrier the other was a lap behind on: the parity
// wait aliases and a stage is overwritten unconsumed -- found the hard way.)
#include "tc_common.cuh"

#ifdef RDSIC_DEBUG
#define RP_TS(cond, off) ((rg.ts && blockIdx.x == 0 && (cond) && lt < 60) ? rg.ts + lt * 16 + (off) : nullptr)
#define RP_NOMATH (rg.dbg & 2)
#define RP_NOSTORE (rg.dbg & 4)
#define RP_NORES (rg.dbg & 8)
#else
#define RP_TS(cond, off) ((long long*)nullptr)
#define RP_NOMATH 0
#define RP_NOSTORE 0
#define RP_NORES 0
#endif

namespace {

constexpr int RP_EPI_WARPS = 12;
constexpr int RP_TAIL_WARP = 2 + RP_EPI_WARPS;          // issues the tail GEMMs (leader CTA)
constexpr int RP_THREADS = 128 + 32 * RP_EPI_WARPS;
constexpr int RP_PARTS = RP_EPI_WARPS / 4;
constexpr int RP_MAXC = 96;                              // 5 C <= 512 TMEM columns
constexpr int RP_CHUNKS1 = RP_MAXC / 16 / RP_PARTS;      // phase-1 chunks per warp (2)
constexpr int RP_CHUNKS2 = 2 * RP_MAXC / 16 / RP_PARTS;  // phase-2 chunks per warp (4)
constexpr int RP_SLOTS = 2;       // one patch slot per channel block (the main loop runs far ahead of the epilogue anyway)
constexpr int RP_SLOT_SHIFT = 1;  // log2(RP_SLOTS)
constexpr int RP_RES_WARP = RP_TAIL_WARP + 1;  // TMA producer of the residual tiles
constexpr int RP_RES_BLOCK = BM * BK * 2;      // one 64-channel block of a residual tile: 128 rows x 128 B

struct RuPairGeom {
  int kiters, kb, kc_last, k2_blocks, kc2_last, N2;
  int halo_w, halo_h, patch_bytes, patch_tx;  // (TW+KW-1) x (TH+KH-1) pixels x 128 B; patch_bytes rounded up to 1 KB
  long long* ts;  // RDSIC_DEBUG builds (RDSIC_RP_TS = device address of an int64 buffer): clock64 stamps of CTA 0, 16 per local tile
  int dbg;        // RDSIC_DEBUG builds (RDSIC_RP_DBG): 2 = no epilogue math / stores, 4 = no stores, 8 = no residual loads (garbage results)
  int b_blk_bytes, w3_blk_bytes;  // one resident k-block of this CTA's half: (C/2) x 128 B, (N2/2) x 128 B
};

__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void rp_tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void rp_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__global__ void __launch_bounds__(RP_THREADS, 1)
ru_pair_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_r, const rdsic_conv_desc d,
                  const TcGeom g, const RuPairGeom rg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  // Every byte of the 227 KB is spoken for, so there is no slack for re-aligning the window: the dynamic window of a
  // kernel without static shared memory starts 1 KB into the SM's shared memory, i.e. 1024-aligned (checked below).
  uint8_t* smem = smem_raw;
  uint8_t* b_res = smem + (size_t)RP_SLOTS * rg.patch_bytes;         // [kiters][C/2 rows x 128 B]
  uint8_t* w3_res = b_res + (size_t)rg.kiters * rg.b_blk_bytes;      // [k2_blocks][N2/2 rows x 128 B]
  uint8_t* res_s = w3_res + (size_t)rg.k2_blocks * rg.w3_blk_bytes;  // residual tile: N2/64 blocks of [128 rows x 128 B]
  uint64_t* a_full = (uint64_t*)(res_s + (size_t)(rg.N2 / BK) * RP_RES_BLOCK);      // [RP_SLOTS] LEADER's copy
  uint64_t* a_empty = a_full + RP_SLOTS;         // [RP_SLOTS] local, multicast commit of the issuer
  uint64_t* acc1_full = a_empty + RP_SLOTS;      // [2] local, multicast commit of the issuer
  uint64_t* p_full = acc1_full + 2;              // [2] LEADER's copy: P[b] staged and acc1[b] drained by both CTAs
  uint64_t* acc2_full = p_full + 2;              // local, multicast commit of the tail issuer
  uint64_t* acc2_empty = acc2_full + 1;          // LEADER's copy: acc2 drained by both CTAs
  uint64_t* w_full = acc2_empty + 1;             // LEADER's copy: resident weights of both CTAs have landed
  uint64_t* r_full = w_full + 1;                 // local: this CTA's residual tile has landed
  uint64_t* r_empty = r_full + 1;                // local: ... and has been read by the twelve epilogue warps
  uint32_t* tmem_slot = (uint32_t*)(r_empty + 1);
  float* bias2_s = (float*)(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);  // [N2] tail bias; b2 is read through L1

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int C = d.Cout;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_g) : "memory");
    for (int s = 0; s < RP_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&acc1_full[k], 1);
      mbar_init(&p_full[k], 2 * RP_EPI_WARPS);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 2 * RP_EPI_WARPS);
    mbar_init(w_full, 1);
    mbar_init(r_full, 1);
    mbar_init(r_empty, RP_EPI_WARPS);
    if (smem_u32(smem_raw) & 1023u) __trap();  // (see the carve-up above)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of both CTAs: one allocation for the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  pdl_wait();  // PDL: barrier init / TMEM allocation above overlap the previous kernel's tail; global memory from here on
  for (int i = threadIdx.x; i < rg.N2; i += blockDim.x) bias2_s[i] = d.tail_bias[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();  // the peer's barriers are initialised before any remote copy / commit / arrive targets them
  const uint32_t tmem_base = *tmem_slot;
  const TileWalk wk = make_walk(g);
  const bool leader = wk.crank == 0;
  // TMEM columns
  const uint32_t P_COL = (uint32_t)(2 * C), ACC2_COL = (uint32_t)(3 * C);

  if (warp == 0) {
    // ================= TMA producer (whole warp, elected lane issues) =================
    const uint32_t patch0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t a_full0 = __shfl_sync(0xffffffffu, smem_u32(a_full), 0);
    const uint32_t a_empty0 = __shfl_sync(0xffffffffu, smem_u32(a_empty), 0);
    const uint32_t lead_a_full0 = mapa_u32(a_full0, 0u);
    int kb = rg.kb, kiters = rg.kiters, total = g.walk_total, step = wk.step, KW = d.KW, Cin = d.Cin;
    asm volatile("" : "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step), "+r"(KW), "+r"(Cin));
    {
      // resident weights, once: this CTA's half of the rows of every k-block; all bytes complete on the leader's w_full
      const uint32_t wbar = smem_u32(w_full), lead_wbar = mapa_u32(wbar, 0u);
      const uint32_t b0 = __shfl_sync(0xffffffffu, smem_u32(b_res), 0), w0 = __shfl_sync(0xffffffffu, smem_u32(w3_res), 0);
      if (elect_one()) {
        if (leader) mbar_expect_tx_u32(wbar, 2u * (uint32_t)(kiters * rg.b_blk_bytes + rg.k2_blocks * rg.w3_blk_bytes));
        int n = 0;
        for (int tap = 0; tap < d.KH * KW; ++tap)
          for (int cb = 0; cb < kb; ++cb, ++n)
            tma_load_2d_2sm_u32(b0 + (uint32_t)(n * rg.b_blk_bytes), &tmap_b, lead_wbar, tap * Cin + cb * BK, wk.crank * (C / 2));
        for (int k2 = 0; k2 < rg.k2_blocks; ++k2)
          tma_load_2d_2sm_u32(w0 + (uint32_t)(k2 * rg.w3_blk_bytes), &tmap_g, lead_wbar, k2 * BK, wk.crank * (rg.N2 / 2));
      }
      __syncwarp();
    }
    // patches in consumption order: u = 2 * local tile + channel block, slot u % RP_SLOTS, phase (u / RP_SLOTS) & 1
    uint32_t u = 0;
    for (int q = wk.first; q < total; q += step) {
      int nt, tx, ty, b;
      tile_of(g, wk, q, nt, tx, ty, b);
      const int x0 = tx * g.TW - d.pad_w, y0 = ty * g.TH - d.pad_h;
      for (int cb = 0; cb < 2; ++cb, ++u) {
        const uint32_t slot = u & (RP_SLOTS - 1), ph = (u >> RP_SLOT_SHIFT) & 1u;
        mbar_wait_u32(a_empty0 + 8u * slot, ph ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx_u32(a_full0 + 8u * slot, 2u * (uint32_t)rg.patch_tx);
          tma_load_4d_2sm_u32(patch0 + slot * (uint32_t)rg.patch_bytes, &tmap_a, lead_a_full0 + 8u * slot, cb * BK, x0, y0, b);
        }
        __syncwarp();
      }
    }
    // drain: no multicast commit of the leader may arrive on this CTA's barriers after it has exited -- wait for the
    // release of the LAST fill of every slot
    for (int n = 0; n < RP_SLOTS; ++n, ++u) mbar_wait_u32(a_empty0 + 8u * (u & (RP_SLOTS - 1)), ((u >> RP_SLOT_SHIFT) & 1u) ^ 1u);
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer: leader CTA only =================
    if (leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_full0 = __shfl_sync(0xffffffffu, smem_u32(a_full), 0);
      const uint32_t a_empty0 = __shfl_sync(0xffffffffu, smem_u32(a_empty), 0);
      const uint32_t patch_u0 = (__shfl_sync(0xffffffffu, smem_u32(smem), 0) & 0x3FFFFu) >> 4, patch_u = (uint32_t)rg.patch_bytes >> 4;
      const uint32_t b_u0 = (__shfl_sync(0xffffffffu, smem_u32(b_res), 0) & 0x3FFFFu) >> 4, bblk_u = (uint32_t)rg.b_blk_bytes >> 4;
      const uint64_t dconst = make_sw128_desc(0);
      const uint64_t dconst_halo = make_sw128_desc_ex(0, (uint32_t)(rg.halo_w * 128), 0);  // 8-pixel row groups one patch row apart
      const uint32_t idesc = make_idesc(C, 2 * BM);
      int kb = rg.kb, total = g.walk_total, step = wk.step, KH = d.KH, KW = d.KW, hw8 = rg.halo_w * 8;
      asm volatile("" : "+r"(kb), "+r"(total), "+r"(step), "+r"(KH), "+r"(KW), "+r"(hw8));
      mbar_wait(w_full, 0);  // both CTAs' resident weights are in place
      tcgen05_fence_after();
      uint32_t lt = 0;
      for (int q = wk.first; q < total; q += step, ++lt) {
        const uint32_t b = lt & 1u, use = (lt >> 1) & 1u;
        long long* tsp = RP_TS(lane == 0, 0);
        if (tsp) tsp[0] = clock64();
        mbar_wait(&p_full[b], use ^ 1u);  // acc1[b] drained by BOTH CTAs (phase 1 of tile lt - 2)
        tcgen05_fence_after();
        if (tsp) tsp[1] = clock64();
        const uint32_t acc = tbase + b * (uint32_t)C;
        for (uint32_t cb = 0; cb < 2; ++cb) {
          const uint32_t u = 2u * lt + cb, slot = u & (RP_SLOTS - 1), ph = (u >> RP_SLOT_SHIFT) & 1u;
          mbar_wait_u32(a_full0 + 8u * slot, ph);  // both CTAs' patches of this channel block have landed
          tcgen05_fence_after();
          const uint64_t da0 = dconst_halo + (uint64_t)(patch_u0 + slot * patch_u);
          const uint64_t db0 = dconst + (uint64_t)(b_u0 + cb * bblk_u);
          const int kc = (int)cb + 1 != kb ? 4 : rg.kc_last;  // valid 16-wide K steps of this channel block
          if (elect_one()) {
            int tap = 0;
            for (int r = 0; r < KH; ++r)
              for (int sx = 0; sx < KW; ++sx, ++tap) {
                const uint64_t da = da0 + (uint64_t)(uint32_t)(r * hw8 + sx * 8);     // window shifted by (r, sx) pixels
                const uint64_t db = db0 + (uint64_t)((uint32_t)(tap * kb) * bblk_u);  // resident k-block tap * kb + cb
                if (kc == 4) {
                  umma_bf16_2sm(acc, da, db, idesc, (cb | (uint32_t)tap) ? 1u : 0u);
                  umma_bf16_2sm(acc, da + 2, db + 2, idesc, 1u);
                  umma_bf16_2sm(acc, da + 4, db + 4, idesc, 1u);
                  umma_bf16_2sm(acc, da + 6, db + 6, idesc, 1u);
                } else {
                  for (int k = 0; k < kc; ++k) umma_bf16_2sm(acc, da + 2 * k, db + 2 * k, idesc, (cb | (uint32_t)tap | (uint32_t)k) ? 1u : 0u);
                }
              }
            tcgen05_commit_2sm_mc_u32(a_empty0 + 8u * slot, 3);  // frees the patch slot in BOTH CTAs
          }
          __syncwarp();
        }
        if (elect_one()) tcgen05_commit_2sm_mc_u32(smem_u32(&acc1_full[b]), 3);
        __syncwarp();
        if (tsp) tsp[2] = clock64();
      }
    }
    __syncwarp();
  } else if (warp == RP_TAIL_WARP) {
    // ================= tail-GEMM issuer: leader CTA only =================
    // tail GEMM of tile lt for BOTH CTAs: A = P[b] (each SM's own TMEM), B = resident W3 halves, D = acc2
    if (leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t w3_addr = __shfl_sync(0xffffffffu, smem_u32(w3_res), 0);
      const uint32_t idesc2 = make_idesc(rg.N2, 2 * BM);
      int total = g.walk_total, step = wk.step;
      asm volatile("" : "+r"(total), "+r"(step));
      mbar_wait(w_full, 0);
      uint32_t lt = 0;
      for (int q = wk.first; q < total; q += step, ++lt) {
        const uint32_t b = lt & 1u, par1 = (lt >> 1) & 1u;
        long long* tsp = RP_TS(lane == 0, 0);
        mbar_wait(&p_full[b], par1);             // all 24 epilogue warps of the pair have staged P[b]
        mbar_wait(acc2_empty, (lt & 1u) ^ 1u);   // ... and drained acc2 (phase 2 of tile lt - 1)
        tcgen05_fence_after();
        if (tsp) tsp[6] = clock64();
        const uint32_t p_t = tbase + P_COL + b * (uint32_t)(C / 2), acc2 = tbase + ACC2_COL;
        if (elect_one()) {
          for (int kb2 = 0; kb2 < rg.k2_blocks; ++kb2) {
            const uint64_t dg = make_sw128_desc(w3_addr + (uint32_t)(kb2 * rg.w3_blk_bytes));
            const int kc2 = kb2 + 1 == rg.k2_blocks ? rg.kc2_last : BK / 16;
            for (int k = 0; k < kc2; ++k)  // 16 bf16 of K = 8 TMEM columns of the staged operand
              umma_bf16_ts_2sm(acc2, p_t + (uint32_t)((kb2 * 4 + k) * 8), dg + 2 * k, idesc2, (kb2 | k) ? 1u : 0u);
          }
          tcgen05_commit_2sm_mc_u32(smem_u32(acc2_full), 3);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == RP_RES_WARP) {
    // ================= residual producer (both CTAs): the tile's x rows as N2/64 swizzled [128 rows x 128 B] blocks
    // (the per-lane 32-byte LDG.256 of the first versions -- 32 sectors per instruction, issued in bursts between
    // compute phases and overlapped with nothing -- cost 39 of 230 us per launch.  Sending the OUTPUT through the same
    // buffer and a TMA store was also built: correct, but 8 us slower than per-lane STG.256 -- store, wait for its
    // shared-memory reads, then the next tile's load serialise on the one buffer there is room for.)
    const uint32_t res0 = __shfl_sync(0xffffffffu, smem_u32(res_s), 0);
    const uint32_t rf = __shfl_sync(0xffffffffu, smem_u32(r_full), 0), re = __shfl_sync(0xffffffffu, smem_u32(r_empty), 0);
    int total = g.walk_total, step = wk.step, nblk = rg.N2 / BK;
    asm volatile("" : "+r"(total), "+r"(step), "+r"(nblk));
    uint32_t lt = 0;
    for (int q = wk.first; q < total; q += step, ++lt) {
      int nt, tx, ty, b;
      tile_of(g, wk, q, nt, tx, ty, b);
      mbar_wait_u32(re, (lt & 1u) ^ 1u);  // phase 2 of the previous tile has read the buffer
      if (elect_one()) {
        mbar_expect_tx_u32(rf, (uint32_t)(nblk * RP_RES_BLOCK));
        for (int k = 0; k < nblk; ++k)
          tma_load_4d_u32(res0 + (uint32_t)(k * RP_RES_BLOCK), &tmap_r, rf, k * BK, tx * g.TW, ty * g.TH, b);
      }
      __syncwarp();
    }
    __syncwarp();
  } else if (warp < RP_TAIL_WARP) {
    // ================= epilogue warps (both CTAs) =================
    const int q = warp % 4, part = (warp - 2) / 4;
    const int ml = q * 32 + lane;
    const int dy = ml / g.TW, dx = ml % g.TW;
    const int nchunks1 = C / 16, nchunks2 = rg.N2 / 16;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t lead_p_full0 = mapa_u32(smem_u32(p_full), 0u), lead_acc2_empty = mapa_u32(smem_u32(acc2_empty), 0u);
    const uint32_t acc1_full0 = smem_u32(acc1_full), acc2_full_a = smem_u32(acc2_full);
    const float4* bias1_g = reinterpret_cast<const float4*>(d.bias);        // 384 + 768 B, L1-resident after the first tile
    const uint32_t bias2_a = smem_u32(bias2_s);
    const uint32_t res_row = smem_u32(res_s) + (uint32_t)(ml * 128), r_full_a = smem_u32(r_full), r_empty_a = smem_u32(r_empty);
    const uint32_t rsw = (uint32_t)(ml & 7);  // 128B swizzle: 16-byte chunk index ^ (row & 7)
    const int total = g.walk_total, step = wk.step;

    // ---- phase 1 of local tile lt: t = gelu(acc1 + b2) -> bf16 -> TMEM operand P[lt & 1]
    auto phase1 = [&](uint32_t lt) {
      const uint32_t bsel = lt & 1u, par1 = (lt >> 1) & 1u;
      const uint32_t acc1_t = tlane + bsel * (uint32_t)C, p_t = tlane + P_COL + bsel * (uint32_t)(C / 2);
      long long* tsp = RP_TS(warp == 2 && lane == 0, 0);
      if (tsp) tsp[7] = clock64();
      mbar_wait_u32(acc1_full0 + 8u * bsel, par1);
      tcgen05_fence_after();
      if (tsp) tsp[8] = clock64();
      uint32_t ua[RP_CHUNKS1][16];
#pragma unroll
      for (int ci = 0; ci < RP_CHUNKS1; ++ci)
        if (part + RP_PARTS * ci < nchunks1) tmem_ld16_issue(acc1_t + (uint32_t)((part + RP_PARTS * ci) * 16), ua[ci]);
      tmem_ld_wait();
#pragma unroll
      for (int ci = 0; ci < RP_CHUNKS1; ++ci) {
        const int j = part + RP_PARTS * ci;
        if (j >= nchunks1) break;
        tmem_ld_fence(ua[ci]);
        uint32_t st[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 f = __ldg(bias1_g + j * 4 + i);
          const float v0 = __uint_as_float(ua[ci][4 * i]) + f.x, v1 = __uint_as_float(ua[ci][4 * i + 1]) + f.y;
          const float v2 = __uint_as_float(ua[ci][4 * i + 2]) + f.z, v3 = __uint_as_float(ua[ci][4 * i + 3]) + f.w;
          __nv_bfloat162 h0 = RP_NOMATH ? __floats2bfloat162_rn(v0, v1) : __floats2bfloat162_rn(gelu_fast(v0), gelu_fast(v1));
          __nv_bfloat162 h1 = RP_NOMATH ? __floats2bfloat162_rn(v2, v3) : __floats2bfloat162_rn(gelu_fast(v2), gelu_fast(v3));
          st[2 * i] = *reinterpret_cast<uint32_t*>(&h0);
          st[2 * i + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        rp_tmem_st8(p_t + (uint32_t)(j * 8), st);
      }
      rp_tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_u32(lead_p_full0 + 8u * bsel);  // P[bsel] staged, acc1[bsel] drained
      if (tsp) tsp[9] = clock64();
    };

    uint32_t lt = 0;
    if (wk.first < total) phase1(0);
    for (int tq = wk.first; tq < total; tq += step, ++lt) {
      int nt, tx, ty, b;
      const bool tile_ok = tile_of(g, wk, tq, nt, tx, ty, b);
      const int oy = ty * g.TH + dy, ox = tx * g.TW + dx;
      const bool row_ok = tile_ok && oy < d.OH && ox < d.OW;
      const size_t pix = ((size_t)b * d.OHt + (oy * d.osy + d.ooy)) * d.OWt + (ox * d.osx + d.oox);
      __nv_bfloat16* outp = (__nv_bfloat16*)d.out.ptr + pix * (size_t)d.out.ld + d.out.coff;
      if (tq + step < total) phase1(lt + 1);

      // ---- phase 2 of tile lt: out = gelu(acc2 + b3 + x)
      long long* tsp = RP_TS(warp == 2 && lane == 0, 0);
      if (tsp) tsp[10] = clock64();
      mbar_wait_u32(acc2_full_a, lt & 1u);
      tcgen05_fence_after();
      mbar_wait_u32(r_full_a, lt & 1u);  // the tile's residual rows are in shared memory
      if (tsp) tsp[11] = clock64();
#pragma unroll
      for (int cp = 0; cp < RP_CHUNKS2; cp += 2) {
        const int ja = part + RP_PARTS * cp, jb = ja + RP_PARTS;
        if (ja >= nchunks2) break;
        const bool has_b = jb < nchunks2;
        uint32_t ua[16], ub[16];
        tmem_ld16_issue(tlane + ACC2_COL + (uint32_t)(ja * 16), ua);
        if (has_b) tmem_ld16_issue(tlane + ACC2_COL + (uint32_t)(jb * 16), ub);
        tmem_ld_wait();
        tmem_ld_fence(ua);
        if (has_b) tmem_ld_fence(ub);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          if (hb == 1 && !has_b) break;
          const int j = hb ? jb : ja;
          const uint32_t* u = hb ? ub : ua;
          if (!row_ok || RP_NOMATH) continue;
          Pack8 o;
          // residual: 16 channels = two 16-byte chunks of this row's 128-byte line in block j / 4
          const uint32_t rb = res_row + (uint32_t)((j >> 2) * RP_RES_BLOCK), c16 = (uint32_t)((j & 3) * 2);
          uint32_t xr[8];
          {
            const float4 r0 = lds128(rb + ((c16 ^ rsw) << 4)), r1 = lds128(rb + (((c16 + 1) ^ rsw) << 4));
            xr[0] = __float_as_uint(r0.x); xr[1] = __float_as_uint(r0.y); xr[2] = __float_as_uint(r0.z); xr[3] = __float_as_uint(r0.w);
            xr[4] = __float_as_uint(r1.x); xr[5] = __float_as_uint(r1.y); xr[6] = __float_as_uint(r1.z); xr[7] = __float_as_uint(r1.w);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = lds128(bias2_a + (uint32_t)((j * 16 + 4 * i) * 4));
            const uint32_t xa = xr[2 * i], xb = xr[2 * i + 1];
            const float v0 = gelu_fast(__uint_as_float(u[4 * i]) + f.x + __uint_as_float(xa << 16));
            const float v1 = gelu_fast(__uint_as_float(u[4 * i + 1]) + f.y + __uint_as_float(xa & 0xFFFF0000u));
            const float v2 = gelu_fast(__uint_as_float(u[4 * i + 2]) + f.z + __uint_as_float(xb << 16));
            const float v3 = gelu_fast(__uint_as_float(u[4 * i + 3]) + f.w + __uint_as_float(xb & 0xFFFF0000u));
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
            o.w[2 * i] = *reinterpret_cast<uint32_t*>(&h0);
            o.w[2 * i + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
          if (!RP_NOSTORE) stg256(outp + j * 16, o);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster_u32(lead_acc2_empty);  // acc2 drained: the next tail GEMM may overwrite it
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(r_empty_a) : "memory");  // this warp has read its residual rows
      }
      if (tsp) tsp[12] = clock64();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees TMEM / exits while the other may still use the pair's resources
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

// ResidualUnit tail on CTA pairs.  Called by rdsic_conv_gdn_forward_bf16 (conv_gdn_bf16.cu) AFTER it has validated the
// descriptor; returns RDSIC_RU_PAIR_SKIP when the layer does not qualify (the caller then takes the 1-CTA kernel).
int rdsic_ru_pair_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream) {
  static const int tune = getenv("RDSIC_RU_PAIR") ? atoi(getenv("RDSIC_RU_PAIR")) : 1;
  const int C = d->Cout, N2 = d->tail_n;
  const int sms = rdsic_sm_count();
  if (!tune || d->tail_mode != 3 || N2 != 2 * C || 5 * C > 512 || C % 16 || C > RP_MAXC || d->KH * d->KW < 2 || sms < 2) return -1;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return RDSIC_E_UNSUPPORTED;
  TcGeom g = {};
  const int B = d->B, H = d->H, W = d->W, OH = d->OH, OW = d->OW;
  if (d->stride != 1 || d->KH > 3 || d->KW > 3 || d->osy != 1 || d->osx != 1 || d->out.dtype != RDSIC_BF16) return -1;
  g.TH = 16;  // 8 consecutive pixels of a patch row = one 1024-byte swizzle atom (8 rows of the A operand)
  g.TW = 8;
  g.tiles_x = ceil_div(OW, g.TW);
  g.tiles_y = ceil_div(OH, g.TH);
  g.BN = C;
  g.n_tiles = 1;
  g.total_tiles = B * g.tiles_y * g.tiles_x;
  g.m_tiles = g.total_tiles;
  if (g.m_tiles < 2) return -1;
  g.pair = 1;
  g.walk_total = ceil_div(g.m_tiles, 2);
  g.kb_per_tap = ceil_div(d->Cin, BK);
  g.num_k_iters = d->KH * d->KW * g.kb_per_tap;
  g.tmem_cols = 512;
  RuPairGeom rg;
  rg.kb = g.kb_per_tap;
  if (rg.kb != 2) return -1;  // one channel block per issuer
  rg.kiters = g.num_k_iters;
  rg.kc_last = (d->Cin - (g.kb_per_tap - 1) * BK) / 16;
  rg.N2 = N2;
  static const int tune_dbg = getenv("RDSIC_RP_DBG") ? atoi(getenv("RDSIC_RP_DBG")) : 0;
  static const long long tune_ts = getenv("RDSIC_RP_TS") ? atoll(getenv("RDSIC_RP_TS")) : 0;
  rg.dbg = tune_dbg;
  rg.ts = (long long*)tune_ts;
  rg.k2_blocks = ceil_div(C, BK);
  rg.kc2_last = (C - (rg.k2_blocks - 1) * BK) / 16;
  rg.b_blk_bytes = (C / 2) * BK * 2;
  rg.w3_blk_bytes = (N2 / 2) * BK * 2;
  if (rg.b_blk_bytes % 1024 || rg.w3_blk_bytes % 1024) return -1;  // swizzle atoms: C/2 a multiple of 8
  rg.halo_w = g.TW + d->KW - 1;
  rg.halo_h = g.TH + d->KH - 1;
  rg.patch_tx = rg.halo_w * rg.halo_h * BK * 2;
  rg.patch_bytes = (rg.patch_tx + 1023) / 1024 * 1024;
  const size_t resident = (size_t)rg.kiters * rg.b_blk_bytes + (size_t)rg.k2_blocks * rg.w3_blk_bytes;
  const size_t fixed = (2 * RP_SLOTS + 10) * 8 + 16 + 2 * RP_MAXC * 4;  // barriers + TMEM slot + tail bias (no alignment slack: see the kernel's carve-up)
  const size_t smem = (size_t)RP_SLOTS * rg.patch_bytes + resident + (size_t)(N2 / BK) * RP_RES_BLOCK + fixed;
  if (smem > 227u * 1024u || N2 % BK || !d->bias || ((uintptr_t)d->bias % 16) || ((uintptr_t)d->tail_bias % 16)) return -1;
  g.num_stages = RP_SLOTS;

  CUtensorMap ta, tb, tg;
  {
    const cuuint64_t ld_b = (cuuint64_t)d->in.ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld_b, ld_b * W, ld_b * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)rg.halo_w, (cuuint32_t)rg.halo_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
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
  if (encode_2d(&tg, d->tail_weight, C, N2, N2 / 2) != CUDA_SUCCESS) return RDSIC_E_ARG;
  CUtensorMap tr;
  {  // residual x: [B, OH, OW, N2] bf16 view, one 8 x 16-pixel x 64-channel box per block of a tile
    const cuuint64_t ld_b = (cuuint64_t)d->res.ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)N2, (cuuint64_t)d->OWt, (cuuint64_t)d->OHt, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld_b, ld_b * d->OWt, ld_b * d->OWt * d->OHt};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)g.TW, (cuuint32_t)g.TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    void* base = (void*)((const __nv_bfloat16*)d->res.ptr + d->res.coff);
    if (d->ooy || d->oox || d->OHt != OH || d->OWt != OW) return -1;
    if (encode(&tr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return RDSIC_E_ARG;
  }

  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool track = dev >= 0 && dev < 16;
  if (!track || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(ru_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    if (track) attr_set[dev] = true;
  }
  const int grid = 2 * g.walk_total < (sms & ~1) ? 2 * g.walk_total : (sms & ~1);
  return rdsic_launch(ru_pair_tc_kernel, dim3((unsigned)grid), RP_THREADS, smem, stream, true, ta, tb, tg, tr, *d, g, rg);
}
