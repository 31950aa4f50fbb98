// bf16 tensor-core implicit-GEMM convolution for sm_100a: TMA -> shared memory ->
// tcgen05.mma (accumulator in TMEM) -> tcgen05.ld epilogue.
//
// Replaces, per launch, one reference nn.Conv2d / ConvTranspose2d phase / nn.Linear /
// GDN 1x1 contraction (WACNN/utils.py:116-134, layers/layers.py:29-43,
// layers/gdn.py:62-75, layers/win_attention.py:91,113) plus its elementwise tail
// (RDSIC_EPI_*), on channels-last bf16 activations.
//
// GEMM view: M = output pixels, N = Cout, K = taps x Cin.
//   * The M tile is a TH x TW patch of output pixels of one image (TH*TW = 128).  For
//     filter tap (r,s) and channel block c0 its A operand is ONE 4-D TMA box
//     {64 ch, TW, TH, 1} of the NHWC input at (c0, ox0*stride-pad+s, oy0*stride-pad+r, b),
//     with TMA element strides = conv stride: im2col, zero padding (out-of-bounds fill),
//     stride-2 sampling, torch.cat (channel offset/ld) and the deconv phase shifts are all
//     folded into TMA coordinates -- nothing is materialised.
//   * B (weights, [Cout][taps*Cin] bf16) is a 2-D TMA box {64, BN}.
//   * Both land in 128B-swizzled K-major tiles that tcgen05.mma consumes directly; the
//     accumulator (128 x BN fp32) lives in TMEM and is read back by 4 epilogue warps.
//   * Warp roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2-5 =
//     epilogue (TMEM lane quarter = warp_id % 4).
#include "tc_common.cuh"

#ifdef RDSIC_DEBUG  // profiling aids (time stamps, operand-skip switches): compiled out of release builds
long long* g_dbg_ts = nullptr;
int g_dbg_host = 0;  // see rdsic_debug_read_ts
#endif

namespace {

// ------------------------------------------------------------------ the kernel
// Persistent, warp-specialised: each CTA walks tiles blockIdx.x, +gridDim.x, ... ; the TMA ring and the
// two TMEM accumulator buffers run continuously across tiles, so the epilogue of tile t overlaps the main
// loop of tile t+1 and the producer prefetches the next tile's operands while the last MMAs retire.
//
// EPI: fused epilogue (compile time, keeps the epilogue's code small enough for the I-cache);
// PLAIN: every output/residual view is channels-last with 16-element-aligned rows (vector path),
//        otherwise the generic scalar addressing (NCHW output, PixelShuffle) is used.
//
// K-split (g.ksplit): issuing a tcgen05.mma costs its thread ~77 cycles whatever N is (measured with
// tests/gpu_issue_trace.py), i.e. ~310 of the ~540 cycles one 64-wide K step takes a single issuer while the
// pipe needs 4 x N/2.  Two issuer warps (warp 1 and ISSUER2_WARP) therefore take alternate k-iterations,
// each into its OWN accumulator (fixed summation order -> results stay bit-reproducible run to run); the
// epilogue adds the two.
constexpr int ISSUER2_WARP = 2 + NUM_EPI_WARPS;
constexpr int PRODUCER2_WARP = ISSUER2_WARP + 1;
constexpr int CONV_THREADS = NUM_THREADS + 64;

template <int EPI, bool PLAIN, bool PAIR>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const rdsic_conv_desc d, const TcGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  // PAIR (cta_group::2) is a template parameter: a kernel that contains cta_group::2 instructions can only be launched
  // in clusters of two, so the 1-CTA instantiations must not contain any; the PAIR instantiation in turn drops the
  // halo / M2 / K-split / MC variants
  const int m_halo = PAIR ? 0 : g.halo, m_m2 = PAIR ? 0 : g.m2, m_ksplit = PAIR ? 0 : g.ksplit, m_mc = PAIR ? 0 : g.mc;
  // PAIR + M2 ("P2"): every CTA of the pair stages a 256-row A box (two 128-row halves), the leader issues TWO
  // cta_group::2 UMMAs per K step (half h of both CTAs into accumulator columns h * BN) against the same B stage: 512
  // rows per pair and weight stage, i.e. half the weight bytes per output row of the plain pair tile
  const int m_p2 = PAIR ? g.m2 : 0;
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = m_halo ? g.b_stage_bytes : g.a_stage_bytes + g.b_stage_bytes;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)g.num_stages * stage_bytes + (m_halo ? 2 * g.a_halo_bytes : 0));
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc_full = empty_bar + MAX_STAGES;   // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;            // [2] epilogue -> MMA
  uint64_t* a_full = acc_empty + 2;              // [2] halo mode: A patch ring
  uint64_t* a_empty = a_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(a_empty + 2);
  // the layer's bias staged once per CTA (zeros when the layer has none); 16-byte aligned for float4 reads
  float* bias_s = (float*)(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);
  // halo mode smem: [num_stages x B tile][2 x A halo patch]; otherwise [num_stages x (A tile | B tile)]
  uint8_t* a_halo = smem + (size_t)g.num_stages * g.b_stage_bytes;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
    for (int s = 0; s < g.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (m_m2 ? 2 : 1) * (m_mc ? 2 : 1));  // M2: both issuers drain every stage; MC: both CTAs' issuers
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&acc_full[k], (m_ksplit || m_m2) ? 2 : 1);
      mbar_init(&acc_empty[k], NUM_EPI_WARPS * (PAIR ? 2 : 1));  // PAIR: the leader's barrier collects both CTAs' epilogue warps
      mbar_init(&a_full[k], 1);
      mbar_init(&a_empty[k], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp allocates TMEM (and later frees it); PAIR: the same warp of both CTAs, as one allocation
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(g.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(g.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  // PDL: everything above touched only this CTA's own state; from here on global memory is read (the bias may have been
  // packed by the immediately preceding kernel), so wait for the prerequisite grids first
  pdl_wait();
  for (int i = threadIdx.x; i < (d.Cout + 15) / 16 * 16; i += blockDim.x) bias_s[i] = (d.bias && i < d.Cout) ? d.bias[i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (m_mc || PAIR) cluster_sync_all();  // the peer's barriers are initialised before any remote copy / commit / arrive targets them
  const uint32_t tmem_base = *tmem_slot;
  const TileWalk wk = make_walk(g);
  const uint32_t b_half_off = (uint32_t)(wk.crank * (g.BN / 2) * BK * 2);  // MC: this CTA's half of the B stage
  const int n_half = wk.crank * (g.BN / 2);                                 // MC / PAIR: first B row this CTA loads

  if (warp == 0 || warp == PRODUCER2_WARP) {
    // ================= TMA producers =================
    // NOTE: no integer division inside the per-stage loops -- this warp's latency paces the whole pipeline.
    if (m_halo) {
      if (lane == 0 && warp == 0) {
        int s = 0, as = 0;
        uint32_t ph = 0, aph = 0;
        const uint32_t a_bytes = (uint32_t)(BK * 2 * g.halo_w * g.halo_h), b_bytes = (uint32_t)g.b_stage_bytes;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
          const int nt = tile % g.n_tiles;
          int t = tile / g.n_tiles;
          const int tx = t % g.tiles_x;
          t /= g.tiles_x;
          const int ty = t % g.tiles_y, b = t / g.tiles_y;
          const int x0 = tx * g.TW * d.stride - d.pad_w, y0 = ty * g.TH * d.stride - d.pad_h, n0 = nt * g.BN;
          for (int cb = 0; cb < g.kb_per_tap; ++cb) {
            mbar_wait(&a_empty[as], aph ^ 1u);
            mbar_expect_tx(&a_full[as], a_bytes);
            tma_load_4d(a_halo + (size_t)as * g.a_halo_bytes, &tmap_a, &a_full[as], cb * BK, x0, y0, b);
            if (++as == 2) { as = 0; aph ^= 1u; }
            int kc0 = cb * BK;
            for (int tap = 0; tap < d.KH * d.KW; ++tap, kc0 += d.Cin) {
              mbar_wait(&empty_bar[s], ph ^ 1u);
              mbar_expect_tx(&full_bar[s], b_bytes);
              tma_load_2d(smem + (size_t)s * stage_bytes, &tmap_b, &full_bar[s], kc0, n0);
              if (++s == g.num_stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    } else {
      // Whole warp, uniform control flow, one elected lane issues (see elect_one).  Two producer warps
      // (warp 0 and PRODUCER2_WARP) share the ring: one warp's wait -> expect_tx -> 2 x TMA chain (~400 cycles)
      // could not feed two MMA issuers.  Ownership is by STAGE parity (the ring depth is even): a stage is
      // always filled by the same producer and, in K-split mode, drained by the same issuer, so every thread
      // meets each mbarrier's phases strictly in order.  (Splitting by k-iteration parity instead lets one
      // thread get two phases ahead of a barrier whose previous use belongs to the other thread; the parity
      // wait then aliases to an old phase and passes early -- seen as sporadic launch failures.)
      const int pw = warp == 0 ? 0 : 1;
      const uint32_t a_bytes = (uint32_t)g.a_stage_bytes;
      const uint32_t tx_bytes = ((g.dbg_skip_load & 1) ? 0u : a_bytes) +
                                ((g.dbg_skip_load & 2) ? 0u : (uint32_t)g.b_stage_bytes);
      const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t full0 = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
      const uint32_t empty0 = __shfl_sync(0xffffffffu, smem_u32(empty_bar), 0);
      const uint32_t lead_full0 = PAIR ? mapa_u32(full0, 0u) : full0;  // PAIR: the leader CTA's full barriers
      int ns = g.num_stages, kb = g.kb_per_tap, kiters = d.KH * d.KW * g.kb_per_tap, total = g.walk_total, step = wk.step;
      int KW = d.KW, Cin = d.Cin;
      asm volatile("" : "+r"(ns), "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step), "+r"(KW), "+r"(Cin));
      const int kq = kiters / ns, kr = kiters % ns;
      int s_base = 0;
      uint32_t ph_base = 0;
      if (m_m2) {
        // M2: one producer walks every stage in ring order (a stage carries a 256-row A box, so the per-stage
        // issue chain is amortised over twice the MMAs; the ring depth may be odd)
        if (pw == 0) {
          int s = 0;
          uint32_t ph = 0;
          for (int q = wk.first; q < total; q += step) {
            int nt, tx, ty, b;
            tile_of(g, wk, q, nt, tx, ty, b);
            const int x0 = tx * g.TW * d.stride - d.pad_w, y0 = ty * g.TH * d.stride - d.pad_h, n0 = nt * g.BN;
            const int ag = nt * g.a_group_stride;  // grouped conv: this n tile's input channels
            int cb = 0, r = 0, sx = 0, kcol = 0;
            for (int n = 0; n < kiters; ++n) {
              mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
              const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes), bar = full0 + 8u * (uint32_t)s;
              if (elect_one()) {
                mbar_expect_tx_u32(bar, tx_bytes);
                if (!(g.dbg_skip_load & 1)) tma_load_4d_u32(a_dst, &tmap_a, bar, ag + cb * BK, x0 + sx, y0 + r, b);
                if (m_mc) tma_load_2d_mc_u32(a_dst + a_bytes + b_half_off, &tmap_b, bar, kcol + cb * BK, n0 + n_half, 3);
                else if (!(g.dbg_skip_load & 2)) tma_load_2d_u32(a_dst + a_bytes, &tmap_b, bar, kcol + cb * BK, n0);
              }
              __syncwarp();
              if (++s == ns) { s = 0; ph ^= 1u; }
              if (++cb == kb) {
                cb = 0;
                kcol += Cin;
                if (++sx == KW) { sx = 0; ++r; }
              }
            }
          }
          // MC drain: leave only when both CTAs have released every stage for the last time -- no multicast commit
          // of the peer may arrive on this CTA's barriers after it has exited
          if (m_mc)
            for (int n = 0; n < ns; ++n) {
              mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
              if (++s == ns) { s = 0; ph ^= 1u; }
            }
        }
      } else
      for (int q = wk.first; q < total; q += step) {
        int nt, tx, ty, b;
        tile_of(g, wk, q, nt, tx, ty, b);
        const int x0 = tx * g.TW * d.stride - d.pad_w, y0 = ty * g.TH * d.stride - d.pad_h, n0 = nt * g.BN;
        const int ag = nt * g.a_group_stride;  // grouped conv: this n tile's input channels
        // this producer's first k-iteration of the tile (the first whose stage has its parity): stage, filter
        // tap and channel block
        const int f = pw ^ (s_base & 1);
        const int n_own = (kiters - f + 1) / 2;
        int s = s_base + f;
        uint32_t ph = ph_base;
        if (s >= ns) { s -= ns; ph ^= 1u; }
        int cb = f % kb, tap0 = f / kb;
        int r = tap0 / KW, sx = tap0 % KW;
        int kcol = tap0 * Cin;  // K coordinate in the packed weight = tap * Cin + cb * 64
        for (int n = 0; n < n_own; ++n) {
          mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
          const uint32_t a_dst = smem_base + (uint32_t)(s * stage_bytes), bar = full0 + 8u * (uint32_t)s;
          if (PAIR) {
            // both CTAs' boxes complete on the LEADER's full barrier, which expects the bytes of both
            const uint32_t lbar = lead_full0 + 8u * (uint32_t)s;
            if (elect_one()) {
              if (wk.crank == 0) mbar_expect_tx_u32(bar, 2u * tx_bytes);
              tma_load_4d_2sm_u32(a_dst, &tmap_a, lbar, ag + cb * BK, x0 + sx, y0 + r, b);
              tma_load_2d_2sm_u32(a_dst + a_bytes, &tmap_b, lbar, kcol + cb * BK, n0 + n_half);
            }
          } else if (elect_one()) {
            mbar_expect_tx_u32(bar, tx_bytes);
            if (!(g.dbg_skip_load & 1)) tma_load_4d_u32(a_dst, &tmap_a, bar, ag + cb * BK, x0 + sx, y0 + r, b);
            if (m_mc) tma_load_2d_mc_u32(a_dst + a_bytes + b_half_off, &tmap_b, bar, kcol + cb * BK, n0 + n_half, 3);
            else if (!(g.dbg_skip_load & 2)) tma_load_2d_u32(a_dst + a_bytes, &tmap_b, bar, kcol + cb * BK, n0);
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
      if ((m_mc || PAIR) && !m_m2) {  // MC / PAIR drain of this producer's stages (see the M2 producer)
        int s = s_base;
        uint32_t ph = ph_base;
        for (int n = 0; n < ns; ++n) {
          if ((s & 1) == pw) mbar_wait_u32(empty0 + 8u * (uint32_t)s, ph ^ 1u);
          if (++s == ns) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == ISSUER2_WARP) {
    // ================= MMA issuers =================
    const uint32_t me = warp == 1 ? 0u : 1u;
    const uint32_t idesc = make_idesc(g.BN);
    const int taps = d.KH * d.KW;
    const int kc_last = (d.Cin - (g.kb_per_tap - 1) * BK) / 16;  // valid 16-wide K steps of a tap's last block
    if (m_halo) {
      if (lane == 0 && !me) {  // (experimental mode, divergent single-lane issue)
        int s = 0, as = 0;
        uint32_t ph = 0, lt = 0, aph = 0;
        for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++lt) {
          const uint32_t buf = lt & 1u, cph = (lt >> 1) & 1u;
          mbar_wait(&acc_empty[buf], cph ^ 1u);
          tcgen05_fence_after();
          const uint32_t acc = tmem_base + buf * (uint32_t)g.BN;
          uint32_t accumulate = 0;
          for (int cb = 0; cb < g.kb_per_tap; ++cb) {
            mbar_wait(&a_full[as], aph);
            tcgen05_fence_after();
            const int kc = cb + 1 == g.kb_per_tap ? kc_last : BK / 16;
            const uint32_t a_base = smem_u32(a_halo + (size_t)as * g.a_halo_bytes);
            int r = 0, sx = 0;
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(&full_bar[s], ph);
              tcgen05_fence_after();
              const uint32_t a_start = a_base + (uint32_t)((r * g.halo_w + sx) * 128);
              const uint64_t da = make_sw128_desc_ex(a_start, (uint32_t)(g.halo_w * 128),
                                                     g.halo_base_off ? (a_start >> 7) & 7u : 0u);
              const uint64_t db = make_sw128_desc(smem_u32(smem + (size_t)s * stage_bytes));
              for (int k = 0; k < kc; ++k) {
                umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
              tcgen05_commit(&empty_bar[s]);
              if (++s == g.num_stages) { s = 0; ph ^= 1u; }
              if (++sx == d.KW) { sx = 0; ++r; }
            }
            tcgen05_commit(&a_empty[as]);  // the patch is free once this block's MMAs retire
            if (++as == 2) { as = 0; aph ^= 1u; }
          }
          tcgen05_commit(&acc_full[buf]);
        }
      }
    } else {
      // Whole warp in uniform control flow, one elected lane issues (see elect_one): every operand of the
      // UTCHMMA then lives in uniform registers.  The loops below are written for a minimal instruction count
      // per k-iteration -- the issuing thread's own instruction stream (~120 instructions per k-iteration in
      // the first version, against 4 x ~77 cycles for the MMAs themselves) paces layers with N <= 128:
      //  * each issuer walks only its own k-iterations (stage index += ways, no skipped iterations),
      //  * descriptors are one 64-bit add from a per-kernel constant (start-address field = smem address >> 4),
      //  * kernel parameters used in the loop are pinned in registers (no constant-bank reloads),
      //  * full 64-wide K blocks take a branch-free 4-MMA path.
      const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t empty0 = __shfl_sync(0xffffffffu, smem_u32(empty_bar), 0);
      const uint32_t full0 = __shfl_sync(0xffffffffu, smem_u32(full_bar), 0);
      const uint64_t dconst = make_sw128_desc(0);                   // everything but the start-address field
      const uint32_t a_u0 = (smem_base & 0x3FFFFu) >> 4, stage_u = (uint32_t)stage_bytes >> 4;
      int ns = g.num_stages, kb = g.kb_per_tap, kiters = taps * g.kb_per_tap, total = g.walk_total, step = wk.step;
      asm volatile("" : "+r"(ns), "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step));
      const bool mc = m_mc != 0;
      const int ways = m_ksplit ? 2 : 1;
      const int kq = kiters / ns, kr = kiters % ns;                  // per-tile advance of the stage ring
      int s_base = 0, dbg_n = 0;
      uint32_t ph_base = 0, lt = 0;
      if (PAIR) {
        // PAIR: the leader CTA's first issuer warp issues every UMMA for both CTAs (256 x BN per K step); K runs
        // sequentially in one accumulator per SM, i.e. in the single-issuer summation order
        if (!me && wk.crank == 0) {
          const uint32_t idesc2 = make_idesc(g.BN, 2 * BM);
          const uint32_t b_off = (uint32_t)g.a_stage_bytes >> 4;
          int s = 0, cb = 0;
          uint32_t ph = 0;
          for (int q = wk.first; q < total; q += step, ++lt) {
            const uint32_t buf = g.acc_bufs == 2 ? (lt & 1u) : 0u, cph = g.acc_bufs == 2 ? ((lt >> 1) & 1u) : (lt & 1u);
            mbar_wait(&acc_empty[buf], cph ^ 1u);  // both CTAs' epilogues have drained this accumulator buffer
            tcgen05_fence_after();
            const uint32_t acc = tbase + buf * (uint32_t)g.acc_stride;
            for (int n = 0; n < kiters; ++n) {
              mbar_wait_u32(full0 + 8u * (uint32_t)s, ph);  // both CTAs' A boxes and B halves have landed
              tcgen05_fence_after();
              const uint64_t da = dconst + (uint64_t)(a_u0 + (uint32_t)s * stage_u), db = da + b_off;
              if (elect_one()) {
                if (cb + 1 != kb || kc_last == 4) {
                  umma_bf16_2sm(acc, da, db, idesc2, n > 0 ? 1u : 0u);
                  umma_bf16_2sm(acc, da + 2, db + 2, idesc2, 1u);
                  umma_bf16_2sm(acc, da + 4, db + 4, idesc2, 1u);
                  umma_bf16_2sm(acc, da + 6, db + 6, idesc2, 1u);
                  if (m_p2) {  // second 128-row half of both CTAs' A boxes, same B stage
                    const uint64_t da2 = da + (A_STAGE_BYTES >> 4);
                    const uint32_t acc2 = acc + (uint32_t)g.BN;
                    umma_bf16_2sm(acc2, da2, db, idesc2, n > 0 ? 1u : 0u);
                    umma_bf16_2sm(acc2, da2 + 2, db + 2, idesc2, 1u);
                    umma_bf16_2sm(acc2, da2 + 4, db + 4, idesc2, 1u);
                    umma_bf16_2sm(acc2, da2 + 6, db + 6, idesc2, 1u);
                  }
                } else {
                  for (int k = 0; k < kc_last; ++k) umma_bf16_2sm(acc, da + 2 * k, db + 2 * k, idesc2, (n > 0 || k > 0) ? 1u : 0u);
                  if (m_p2)
                    for (int k = 0; k < kc_last; ++k)
                      umma_bf16_2sm(acc + (uint32_t)g.BN, da + (A_STAGE_BYTES >> 4) + 2 * k, db + 2 * k, idesc2, (n > 0 || k > 0) ? 1u : 0u);
                }
                tcgen05_commit_2sm_mc_u32(empty0 + 8u * (uint32_t)s, 3);  // frees the stage in BOTH CTAs
              }
              __syncwarp();
              if (++s == ns) { s = 0; ph ^= 1u; }
              if (++cb == kb) cb = 0;
            }
            if (elect_one()) tcgen05_commit_2sm_mc_u32(smem_u32(&acc_full[buf]), 3);
            __syncwarp();
          }
        }
      } else if (m_m2) {
        // M2: both issuers visit EVERY stage in ring order; issuer `me` multiplies rows [128 me, 128 me + 128) of
        // the 256-row A stage with the shared B stage into its own accumulator (columns me * BN).  K runs
        // sequentially in one accumulator, so the fp32 summation order equals the single-issuer path's.
        const uint32_t a_half = me * (uint32_t)(A_STAGE_BYTES >> 4), b_off = (uint32_t)g.a_stage_bytes >> 4;
        int s = 0, cb = 0;
        uint32_t ph = 0;
        for (int q = wk.first; q < total; q += step, ++lt) {
          const uint32_t buf = g.acc_bufs == 2 ? (lt & 1u) : 0u, cph = g.acc_bufs == 2 ? ((lt >> 1) & 1u) : (lt & 1u);
          mbar_wait(&acc_empty[buf], cph ^ 1u);
          tcgen05_fence_after();
          const uint32_t acc = tbase + buf * (uint32_t)g.acc_stride + me * (uint32_t)g.BN;
          for (int n = 0; n < kiters; ++n) {
            mbar_wait_u32(full0 + 8u * (uint32_t)s, ph);
            tcgen05_fence_after();
            const uint64_t st0 = dconst + (uint64_t)(a_u0 + (uint32_t)s * stage_u);
            const uint64_t da = st0 + a_half, db = st0 + b_off;
            if (elect_one()) {
              if (cb + 1 != kb || kc_last == 4) {
                umma_bf16(acc, da, db, idesc, n > 0 ? 1u : 0u);
                umma_bf16(acc, da + 2, db + 2, idesc, 1u);
                umma_bf16(acc, da + 4, db + 4, idesc, 1u);
                umma_bf16(acc, da + 6, db + 6, idesc, 1u);
              } else {
                for (int k = 0; k < kc_last; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (n > 0 || k > 0) ? 1u : 0u);
              }
              if (mc) tcgen05_commit_mc_u32(empty0 + 8u * (uint32_t)s, 3);
              else tcgen05_commit_u32(empty0 + 8u * (uint32_t)s);
            }
            __syncwarp();
            if (++s == ns) { s = 0; ph ^= 1u; }
            if (++cb == kb) cb = 0;
          }
          if (elect_one()) tcgen05_commit(&acc_full[buf]);
          __syncwarp();
        }
      } else if (m_ksplit || !me)
      for (int q = wk.first; q < total; q += step, ++lt) {
        const uint32_t buf = g.acc_bufs == 2 ? (lt & 1u) : 0u, cph = g.acc_bufs == 2 ? ((lt >> 1) & 1u) : (lt & 1u);
        // this issuer's first k-iteration of the tile: K-split ownership is by stage parity (see the producers)
        const int f = m_ksplit ? (int)me ^ (s_base & 1) : 0;
        int s = s_base + f;
        uint32_t ph = ph_base;
        if (s >= ns) { s -= ns; ph ^= 1u; }
        int cb = f % kb;
        const int n_own = (kiters - f + ways - 1) / ways;
        mbar_wait(&acc_empty[buf], cph ^ 1u);  // epilogue has drained this accumulator buffer
        const uint32_t acc = tbase + buf * (uint32_t)g.acc_stride + me * (uint32_t)g.BN;
        if (m_ksplit) {
          // two issuers, alternate k-iterations, own accumulators; while one polls its barrier the other issues
          tcgen05_fence_after();
          for (int n = 0; n < n_own; ++n) {
            const bool ts = g.dbg_ts && !(g.dbg_skip_load & 4) && blockIdx.x == 0 && me == 0 && dbg_n < 1024 && lane == 0;
            if (ts) g.dbg_ts[dbg_n * 4 + 0] = clock64();
            mbar_wait_u32(full0 + 8u * (uint32_t)s, ph);
            tcgen05_fence_after();
            if (ts) g.dbg_ts[dbg_n * 4 + 1] = g.dbg_ts[dbg_n * 4 + 2] = clock64();
            const uint64_t da = dconst + (uint64_t)(a_u0 + (uint32_t)s * stage_u), db = da + (A_STAGE_BYTES >> 4);
            if (elect_one()) {
              if (cb + 1 != kb || kc_last == 4) {  // full block: +32 bytes (2 x 16 B units) per 16 bf16 of K
                umma_bf16(acc, da, db, idesc, n > 0 ? 1u : 0u);
                umma_bf16(acc, da + 2, db + 2, idesc, 1u);
                umma_bf16(acc, da + 4, db + 4, idesc, 1u);
                umma_bf16(acc, da + 6, db + 6, idesc, 1u);
              } else {
                for (int k = 0; k < kc_last; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (n > 0 || k > 0) ? 1u : 0u);
              }
              if (mc) tcgen05_commit_mc_u32(empty0 + 8u * (uint32_t)s, 3);
              else tcgen05_commit_u32(empty0 + 8u * (uint32_t)s);
            }
            __syncwarp();
            if (ts) g.dbg_ts[dbg_n++ * 4 + 3] = clock64();
            s += 2;
            if (s >= ns) { s -= ns; ph ^= 1u; }
            cb += 2;
            while (cb >= kb) cb -= kb;
          }
        } else {
          // single issuer, software-pipelined: the barrier wait for the NEXT stage sits between the two halves
          // of this stage's MMAs, so the tensor pipe has work queued while this thread polls the mbarrier
          mbar_wait_u32(full0 + 8u * (uint32_t)s, ph);  // first stage of the tile
          tcgen05_fence_after();
          for (int n = 0; n < kiters; ++n) {
            const bool ts = g.dbg_ts && !(g.dbg_skip_load & 4) && blockIdx.x == 0 && dbg_n < 1024 && lane == 0;
            if (ts) g.dbg_ts[dbg_n * 4 + 0] = clock64();
            const int kc = cb + 1 == kb ? kc_last : BK / 16;
            const uint64_t da = dconst + (uint64_t)(a_u0 + (uint32_t)s * stage_u), db = da + (A_STAGE_BYTES >> 4);
            const uint32_t ebar = empty0 + 8u * (uint32_t)s;
            if (elect_one()) {
              umma_bf16(acc, da, db, idesc, n > 0 ? 1u : 0u);
              if (kc > 1) umma_bf16(acc, da + 2, db + 2, idesc, 1u);
            }
            __syncwarp();
            if (ts) g.dbg_ts[dbg_n * 4 + 1] = clock64();
            if (++s == ns) { s = 0; ph ^= 1u; }
            if (n + 1 < kiters) {
              mbar_wait_u32(full0 + 8u * (uint32_t)s, ph);
              tcgen05_fence_after();
            }
            if (ts) g.dbg_ts[dbg_n * 4 + 2] = clock64();
            if (elect_one()) {
              if (kc > 2) umma_bf16(acc, da + 4, db + 4, idesc, 1u);
              if (kc > 3) umma_bf16(acc, da + 6, db + 6, idesc, 1u);
              if (mc) tcgen05_commit_mc_u32(ebar, 3);
              else tcgen05_commit_u32(ebar);  // frees the smem slot when these MMAs retire
            }
            __syncwarp();
            if (ts) g.dbg_ts[dbg_n++ * 4 + 3] = clock64();
            if (++cb == kb) cb = 0;
          }
        }
        if (elect_one()) tcgen05_commit(&acc_full[buf]);  // (this issuer's) accumulator complete
        __syncwarp();
        // stage ring position of the next tile's first k-iteration
        ph_base ^= (uint32_t)(kq & 1);
        s_base += kr;
        if (s_base >= ns) { s_base -= ns; ph_base ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp < ISSUER2_WARP) {
    // ================= epilogue (warps 2..): EPI_PARTS warps per TMEM lane quarter, interleaved column chunks
    constexpr bool NEED_RES = EPI == RDSIC_EPI_RES_GELU || EPI == RDSIC_EPI_ADD_RES || EPI == RDSIC_EPI_GATE ||
                              EPI == RDSIC_EPI_GDN || EPI == RDSIC_EPI_IGDN || EPI == RDSIC_EPI_LRP;
    constexpr bool NEED_AUX = EPI == RDSIC_EPI_GATE;
    const int q = warp % 4;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) / 4;     // which of the EPI_PARTS warps sharing the quarter
    const int ml = q * 32 + lane;
    // row -> pixel of the patch; M2 tiles have a second 128-row half (rows ml + 128, accumulator columns + BN)
    const int dyh[2] = {ml / g.TW, (ml + BM) / g.TW}, dxh[2] = {ml % g.TW, (ml + BM) % g.TW};
    const int nh = (m_m2 || m_p2) ? 2 : 1;
    const int nchunks = g.BN / 16;
    uint32_t lt = 0;
    for (int tq = wk.first; tq < g.walk_total; tq += wk.step, ++lt) {
      int nt, tx, ty, b;
      const bool tile_ok = tile_of(g, wk, tq, nt, tx, ty, b);
      const int n0 = nt * g.BN;
      const uint32_t buf = g.acc_bufs == 2 ? (lt & 1u) : 0u, aph = g.acc_bufs == 2 ? ((lt >> 1) & 1u) : (lt & 1u);
      bool waited = false;
      for (int h = 0; h < nh; ++h) {
      const int oy = ty * g.TH + dyh[h], ox = tx * g.TW + dxh[h];
      const bool row_ok = tile_ok && oy < d.OH && ox < d.OW;
      const size_t pix = ((size_t)b * d.OHt + (oy * d.osy + d.ooy)) * d.OWt + (ox * d.osx + d.oox);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)g.acc_stride + (uint32_t)(h * g.BN);

      if (PLAIN) {
        // Residual / gate operands are prefetched in groups of G chunks BEFORE the accumulator is awaited,
        // so their HBM latency overlaps this tile's main loop instead of serialising per chunk.
        constexpr int G = NEED_AUX ? 2 : 4;  // register budget: G x (res [+ aux]) 32-byte packs in flight
        const bool res16 = NEED_RES && d.res.dtype == RDSIC_BF16, aux16 = NEED_AUX && d.aux.dtype == RDSIC_BF16;
        const __nv_bfloat16* resp = (const __nv_bfloat16*)d.res.ptr + pix * (size_t)d.res.ld + d.res.coff + n0;
        const __nv_bfloat16* auxp = (const __nv_bfloat16*)d.aux.ptr + pix * (size_t)d.aux.ld + d.aux.coff + n0;
        Pack8 rr[G], ra[G];
        for (int j0 = half; j0 < nchunks; j0 += EPI_PARTS * G) {
#pragma unroll
          for (int gI = 0; gI < G; ++gI) {
            const int j = j0 + EPI_PARTS * gI;
            if (j < nchunks && row_ok && n0 + j * 16 < d.Cout) {
              if (res16) rr[gI] = ldg256(resp + j * 16);
              if (aux16) ra[gI] = ldg256(auxp + j * 16);
            }
          }
          if (!waited) {
            mbar_wait(&acc_full[buf], aph);
            tcgen05_fence_after();
            waited = true;
          }
          // chunks are processed in pairs: both TMEM loads are issued before one wait, and the two
          // independent 16-element epilogues interleave (the epilogue is latency-, not issue-bound)
#pragma unroll
          for (int gI = 0; gI < G; gI += 2) {
            const int ja = j0 + EPI_PARTS * gI, jb = ja + EPI_PARTS;
            if (ja >= nchunks) break;
            const bool has_b = (gI + 1 < G) && jb < nchunks;
            uint32_t ua[16], ub[16];
            tmem_ld16_issue(trow + (uint32_t)(ja * 16), ua);
            if (has_b) tmem_ld16_issue(trow + (uint32_t)(jb * 16), ub);
            if (m_ksplit) {  // add the second issuer's accumulator (fixed order: even + odd k-iterations)
              uint32_t uc[16];
              tmem_ld16_issue(trow + (uint32_t)(g.BN + ja * 16), uc);
              tmem_ld_wait();
              tmem_ld_fence(ua);
              if (has_b) tmem_ld_fence(ub);
              tmem_ld_fence(uc);
#pragma unroll
              for (int i = 0; i < 16; ++i) ua[i] = __float_as_uint(__uint_as_float(ua[i]) + __uint_as_float(uc[i]));
              if (has_b) {
                tmem_ld16_issue(trow + (uint32_t)(g.BN + jb * 16), uc);
                tmem_ld_wait();
                tmem_ld_fence(uc);
#pragma unroll
                for (int i = 0; i < 16; ++i) ub[i] = __float_as_uint(__uint_as_float(ub[i]) + __uint_as_float(uc[i]));
              }
            } else {
              tmem_ld_wait();
              tmem_ld_fence(ua);
              if (has_b) tmem_ld_fence(ub);
            }
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              if (hb == 1 && !has_b) break;
              const int j = hb ? jb : ja, gg = gI + hb;
              const uint32_t* u = hb ? ub : ua;
              const int nb = n0 + j * 16;
              if (!row_ok || nb >= d.Cout) continue;
              float v[16], res[16], aux[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
              if (NEED_RES) {
                if (res16) unpack_bf16x16(rr[gg < G ? gg : 0], res);
                else load16(d.res, pix * (size_t)d.res.ld + d.res.coff + nb, res);
              }
              if (NEED_AUX) {
                if (aux16) unpack_bf16x16(ra[gg < G ? gg : 0], aux);
                else load16(d.aux, pix * (size_t)d.aux.ld + d.aux.coff + nb, aux);
              }
              if (d.bias) {
                const uint32_t bp = smem_u32(bias_s + nb);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 f = lds128(bp + 16u * i);
                  v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = epi_apply<EPI>(v[i], NEED_RES ? res[i] : 0.f, NEED_AUX ? aux[i] : 0.f);
              if (d.pixel_shuffle == 3) {  // columns are sub-position-major: this chunk is 16 channels of ONE shuffled pixel
                const int Cq = d.Cout >> 2, sp = nb / Cq, cq = nb - sp * Cq;
                const size_t px = ((size_t)b * d.OHt + (2 * oy + (sp >> 1))) * d.OWt + (2 * ox + (sp & 1));
                store16(d.out, px * (size_t)d.out.ld + d.out.coff + cq, v, false);
                continue;
              }
              store16(d.out, pix * (size_t)d.out.ld + d.out.coff + nb, v, false);
              if (d.out2.ptr) store16(d.out2, pix * (size_t)d.out2.ld + d.out2.coff + nb, v, d.out2_square != 0);
              if (d.out3.ptr) store16(d.out3, pix * (size_t)d.out3.ld + d.out3.coff + nb, v, false);
            }
          }
        }
        if (!waited) {  // (nchunks <= half cannot happen for BN >= 32, kept for safety)
          mbar_wait(&acc_full[buf], aph);
          tcgen05_fence_after();
          waited = true;
        }
      } else {
        if (!waited) {
          mbar_wait(&acc_full[buf], aph);
          tcgen05_fence_after();
          waited = true;
        }
        for (int j = half; j < nchunks; j += EPI_PARTS) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(j * 16), v);
          if (m_ksplit) {
            float w[16];
            tmem_ld16(trow + (uint32_t)(g.BN + j * 16), w);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w[i];
          }
          const int nb = n0 + j * 16;
          if (!row_ok || nb >= d.Cout) continue;
          const int HWt = d.OHt * d.OWt;
          const int Cview = d.pixel_shuffle ? d.Cout / 4 : d.Cout;
          for (int i = 0; i < 16; ++i) {  // deliberately not unrolled: rare path, keep the code small
            const int n = nb + i;
            if (n >= d.Cout) break;
            float val = v[i] + bias_s[n];
            int c = n;
            size_t px = pix;
            if (d.pixel_shuffle) {
              c = n >> 2;
              px = ((size_t)b * d.OHt + (2 * oy + ((n >> 1) & 1))) * d.OWt + (2 * ox + (n & 1));
            }
            float res = 0.f, aux = 0.f;
            if (NEED_RES) res = ld_elem(d.res.ptr, d.res.dtype, view_index(d.res, px, c, HWt, Cview));
            if (NEED_AUX) aux = ld_elem(d.aux.ptr, d.aux.dtype, view_index(d.aux, px, c, HWt, Cview));
            val = epi_apply<EPI>(val, res, aux);
            st_elem(d.out.ptr, d.out.dtype, view_index(d.out, px, c, HWt, Cview), val);
            if (d.out2.ptr)
              st_elem(d.out2.ptr, d.out2.dtype, view_index(d.out2, px, c, HWt, Cview), d.out2_square ? val * val : val);
            if (d.out3.ptr) st_elem(d.out3.ptr, d.out3.dtype, view_index(d.out3, px, c, HWt, Cview), val);
          }
        }
      }
      }  // h
      // this warp's TMEM reads of the buffer are complete (tcgen05.wait::ld inside tmem_ld16): hand it back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster_u32(mapa_u32(smem_u32(&acc_empty[buf]), 0u));  // the leader's barrier
        else mbar_arrive(&acc_empty[buf]);
      }
    }
  }

  // ---- teardown
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA frees TMEM / exits while the other may still use the pair's resources
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols));
  }
}

typedef void (*ConvTcKernel)(const CUtensorMap, const CUtensorMap, const rdsic_conv_desc, const TcGeom);

template <bool PLAIN, bool PAIR>
ConvTcKernel pick_kernel(int epi) {
  switch (epi) {
    case RDSIC_EPI_GELU: return conv_tc_kernel<RDSIC_EPI_GELU, PLAIN, PAIR>;
    case RDSIC_EPI_RES_GELU: return conv_tc_kernel<RDSIC_EPI_RES_GELU, PLAIN, PAIR>;
    case RDSIC_EPI_ADD_RES: return conv_tc_kernel<RDSIC_EPI_ADD_RES, PLAIN, PAIR>;
    case RDSIC_EPI_GATE: return conv_tc_kernel<RDSIC_EPI_GATE, PLAIN, PAIR>;
    case RDSIC_EPI_GDN: return conv_tc_kernel<RDSIC_EPI_GDN, PLAIN, PAIR>;
    case RDSIC_EPI_IGDN: return conv_tc_kernel<RDSIC_EPI_IGDN, PLAIN, PAIR>;
    case RDSIC_EPI_LRP: return conv_tc_kernel<RDSIC_EPI_LRP, PLAIN, PAIR>;
    default: return conv_tc_kernel<RDSIC_EPI_NONE, PLAIN, PAIR>;
  }
}

// ------------------------------------------------------------------ host side
int pick_bn(int cout) {
  // largest multiple of 16 <= 256 that tiles ceil16(Cout) without remainder (so no wasted MMA columns)
  const int c16 = (cout + 15) / 16 * 16;
  if (c16 <= 256) return c16;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (c16 % bn == 0) return bn;
  return 128;
}

}  // namespace

int rdsic_conv_validate(const rdsic_conv_desc* d);
int rdsic_halo_pair_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream);  // halo_pair_bf16.cu

int rdsic_conv_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream) {
  int rc = rdsic_conv_validate(d);
  if (rc) return rc;
  // tensor-core path requirements (the host packs weights / lays out activations accordingly)
  RDSIC_CHECK_ARG(d->in.dtype == RDSIC_BF16 && !d->in.nchw && !d->a_square);
  RDSIC_CHECK_ARG(d->Cin % 16 == 0);
  if (d->in.ld % 8 || d->in.coff % 8 || ((uintptr_t)d->in.ptr % 16) || ((uintptr_t)d->weight % 16)) return RDSIC_E_ALIGN;
  RDSIC_CHECK_ARG(d->stride == 1 || d->stride == 2);
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return RDSIC_E_UNSUPPORTED;
  rc = rdsic_halo_pair_forward_bf16(d, stream);  // narrow image head on big maps: input read once (halo patches)
  if (rc != -1) return rc;

  TcGeom g = {};
  int B = d->B, H = d->H, W = d->W, OH = d->OH, OW = d->OW;
  // pointwise GEMMs over plain NHWC tensors: flatten all pixels into one row of tiles
  const bool flat = d->KH == 1 && d->KW == 1 && d->stride == 1 && d->pad_h == 0 && d->pad_w == 0 && !d->pixel_shuffle &&
                    !d->out.nchw && !d->res.nchw && !d->aux.nchw && d->osy == 1 && d->osx == 1 && d->ooy == 0 &&
                    d->oox == 0 && OH == H && OW == W && d->OHt == OH && d->OWt == OW;
  rdsic_conv_desc dd = *d;
  if (flat) {
    W = OW = B * H * W;
    H = OH = 1;
    B = 1;
    dd.B = 1; dd.H = 1; dd.W = W; dd.OH = 1; dd.OW = OW; dd.OHt = 1; dd.OWt = OW;
  }
  // patch shape TH x TW (TH*TW = rows): minimise the covered area (= wasted rows), prefer wide patches;
  // a TMA box dimension holds at most 256 elements (TW * stride, TH * stride)
  auto pick_patch = [&](int rows, int* TH, int* TW) {
    if (flat) { *TH = 1; *TW = rows; return (long)ceil_div(OW, rows) * rows; }
    long best = -1;
    for (int tw = rows; tw >= 1; tw /= 2) {
      const int th = rows / tw;
      if (tw * d->stride > 256 || th * d->stride > 256) continue;
      const long area = (long)ceil_div(OW, tw) * tw * ceil_div(OH, th) * th;
      if (best < 0 || area < best) { best = area; *TW = tw; *TH = th; }
    }
    return best;
  };
  pick_patch(BM, &g.TH, &g.TW);
  int dev = 0;
  cudaGetDevice(&dev);
  const bool track = dev >= 0 && dev < 16;
  static int sm_count[16] = {};
  int sms = track ? sm_count[dev] : 0;
  if (!sms) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (track) sm_count[dev] = sms;
  }
  // halo mode: stride-1 multi-tap convs on maps that a 16 x 8 patch tiles exactly (RDSIC_TC_HALO: 0 off,
  // 1 on, 2 on with the descriptor base-offset field derived from the window start)
  static const int tune_halo = getenv("RDSIC_TC_HALO") ? atoi(getenv("RDSIC_TC_HALO")) : 0;
#ifdef RDSIC_DEBUG
  static long long* dbg_ts_buf = nullptr;
  if (getenv("RDSIC_TC_DBG_TS") && !dbg_ts_buf) {
    if (atoi(getenv("RDSIC_TC_DBG_TS")) == 2) {  // host-mapped: survives a trap (timeout log of mbar_wait, tc_common.cuh)
      cudaHostAlloc(&dbg_ts_buf, 16384 * sizeof(long long), cudaHostAllocMapped);
      memset(dbg_ts_buf, 0, 16384 * sizeof(long long));
      g_dbg_host = 1;
      tc_set_timeout_log(dbg_ts_buf);
    } else {
      cudaMalloc(&dbg_ts_buf, 4096 * sizeof(long long));
      cudaMemset(dbg_ts_buf, 0, 4096 * sizeof(long long));
    }
    g_dbg_ts = dbg_ts_buf;
  }
  g.dbg_ts = dbg_ts_buf;
  static const int dbg_skip = getenv("RDSIC_TC_DBG_SKIP") ? atoi(getenv("RDSIC_TC_DBG_SKIP")) : 0;
  g.dbg_skip_load = dbg_skip | (g_dbg_host ? 4 : 0);  // bit 2: the buffer is the timeout log, no time stamps
#endif

  if (tune_halo && d->groups <= 1 && !flat && d->stride == 1 && d->KH * d->KW > 1 && d->KH <= 3 && d->KW <= 3 && OH % 16 == 0 && OW % 8 == 0) {
    g.halo = 1;
    g.halo_base_off = tune_halo == 2;
    g.TH = 16;
    g.TW = 8;
    g.halo_w = g.TW + d->KW - 1;
    g.halo_h = g.TH + d->KH - 1;
    g.a_halo_bytes = (BK * 2 * g.halo_w * g.halo_h + 1023) / 1024 * 1024;
  }
  const int groups = d->groups > 1 ? d->groups : 1;
  if (groups > 1) {  // grouped form: one n tile per group
    RDSIC_CHECK_ARG((d->Cout / groups) % 16 == 0 && d->Cout / groups <= 256 && d->in_group_stride % 8 == 0 && !g.halo);
    g.a_group_stride = d->in_group_stride;
  }
  g.BN = groups > 1 ? d->Cout / groups : pick_bn(d->Cout);
  g.kb_per_tap = ceil_div(d->Cin, BK);
  g.num_k_iters = d->KH * d->KW * g.kb_per_tap;
  g.a_stage_bytes = A_STAGE_BYTES;
  // "Two-issuer family": layers with enough k-iterations to amortise a second issuer warp, decided from layer
  // properties only (Cout, K).  Within the family the mode depends on the grid:
  //   * more 128-row tiles than SMs -> M2: 256-row tiles, the issuers own one 128-row half each and share every
  //     B stage (the big layers are bound by the ~6300 B/cycle chip-wide L2->SM throughput, and the weights
  //     are more than half of the bytes a 128-row tile streams);
  //   * otherwise 128-row tiles with K-split (RDSIC_TC_M2=1) or one sequential issuer (RDSIC_TC_M2=3, which
  //     makes every output bit independent of the batch / image size: M2 and the sequential issuer sum K in
  //     the same order, K-split does not).
  static const int tune_m2 = getenv("RDSIC_TC_M2") ? atoi(getenv("RDSIC_TC_M2")) : 3;
  static const int tune_ksplit = getenv("RDSIC_TC_KSPLIT") ? atoi(getenv("RDSIC_TC_KSPLIT")) : 1;
  static const int tune_mink = getenv("RDSIC_TC_M2_MINK") ? atoi(getenv("RDSIC_TC_M2_MINK")) : 4;
  // PAIR mode (cta_group::2, see TcGeom) replaces M2 / K-split on the wide, long-K layers.  Measured per layer at
  // batch 24 (profiles/r2_pair_vs_m2.txt): N = 224 3x3 layers -5 ... -21 %, 5x5 s2 N = 192 -10 %, N = 1152 (BN 192)
  // -17 %; but N <= 128 and the pointwise layers are 10-60 % SLOWER in pairs: there one k-iteration's tensor time
  // (2 BN cycles) is below what the single issuing thread needs per k-iteration, and M2 / per-SM issuers give each
  // SM its own issue stream.  Hence: pairs for BN >= 192 with at least 8 k-iterations.  RDSIC_TC_PAIR: 0 off,
  // 1 this rule (default), 3 every layer with two M tiles.
  static const int tune_pair = getenv("RDSIC_TC_PAIR") ? atoi(getenv("RDSIC_TC_PAIR")) : 1;
  const long tiles128_all = (long)B * ceil_div(OW, g.TW) * ceil_div(OH, g.TH);
  // (round 2, after the remote arrivals lost their cluster-scope release: the big pointwise layers -- RU heads, qkv,
  // proj, gate at H/4 -- gain ~8 % in pairs, N = 160 / 176 long-K layers gain too; BN <= 128 still loses)
  const bool want_pair = tune_pair && !g.halo && sms >= 2 && tiles128_all >= 2 &&
                         (tune_pair == 3 || (g.BN >= 160 && g.num_k_iters >= 8) ||
                          (tune_pair != 2 && d->KH * d->KW == 1 && g.BN >= 96 && tiles128_all >= 8L * sms));
  const bool family = !g.halo && !want_pair && g.num_k_iters >= tune_mink &&
                      (4 * g.BN <= 512 || (tune_ksplit != 3 && 2 * g.BN <= 512 && g.num_k_iters >= 16));
  if (tune_m2 && family) {
    // cost of the busiest SM ~ rounds x bytes one tile streams per k-iteration (A rows + the B stage)
    const int n_t = ceil_div(d->Cout, g.BN), b_bytes = g.BN * BK * 2;
    const long tiles128 = (long)B * ceil_div(OW, g.TW) * ceil_div(OH, g.TH) * n_t;
    int th2 = 0, tw2 = 0;
    const long area2 = pick_patch(2 * BM, &th2, &tw2);
    const long tiles256 = area2 > 0 ? (long)B * (area2 / (2 * BM)) * n_t : 0;
    const long cost128 = ((tiles128 + sms - 1) / sms) * (A_STAGE_BYTES + b_bytes);
    const long cost256 = ((tiles256 + sms - 1) / sms) * (2 * A_STAGE_BYTES + b_bytes);
    if (area2 > 0 && (tune_m2 == 2 || (tiles128 > sms && cost256 <= cost128))) {
      g.m2 = 1;
      g.TH = th2;
      g.TW = tw2;
      g.a_stage_bytes = 2 * A_STAGE_BYTES;
    }
  }
  // PAIR + M2 (see the kernel): for the pair layers whose grid is a few rounds of pair tiles -- the latent-resolution
  // slice loop: 144 pair tiles on 74 pairs stream the whole weight matrix twice per SM -- a 256-row box per CTA halves
  // the rounds and the weight bytes per output row.  Needs two accumulators of BN columns (one set when 4 BN > 512: the
  // epilogue of a tile is then not hidden behind the next tile's main loop, hence only long-K layers).
  // RDSIC_TC_P2: 0 off, 1 this rule (default), 2 wherever it fits, 3 the rule without the several-n-tiles condition.
  static const int tune_p2 = getenv("RDSIC_TC_P2") ? atoi(getenv("RDSIC_TC_P2")) : 1;
  if (tune_p2 && want_pair && !g.halo && 2 * g.BN <= 512 && groups == 1) {
    const int n_t = ceil_div(d->Cout, g.BN), slots = sms / 2;
    int th2 = 0, tw2 = 0;
    const long area2 = pick_patch(2 * BM, &th2, &tw2);
    const long tiles256 = area2 > 0 ? (long)B * (area2 / (2 * BM)) : 0;
    const long rounds128 = ((tiles128_all + 1) / 2 * n_t + slots - 1) / slots, rounds256 = ((tiles256 + 1) / 2 * n_t + slots - 1) / slots;
    const long cost128 = rounds128 * (A_STAGE_BYTES + g.BN * BK), cost256 = rounds256 * (2 * A_STAGE_BYTES + g.BN * BK);
    // MEASURED per layer at batch 24 (bench.py --dump-ops, RDSIC_TC_P2 = 0 / 2): N = 320 (two n tiles of 160; 3x3 and
    // 5x5 s2 at M = 36 864) -9 ... -12 %; but single-n-tile layers LOSE (N = 224: +9 ... +27 %, N = 192: +5 ... +9 %, the
    // M = 147 456 5x5 s2 layer +7 %): their grid becomes ONE round of 512-row tiles whose whole epilogue is exposed, while
    // two rounds of 256-row tiles hide the first epilogue behind the second main loop.  Hence: at least two rounds of
    // 512-row tiles (several n tiles), where the halved weight traffic is what is left.
    const bool rule = tiles128_all <= 8L * sms && n_t >= 2 && rounds256 >= 2 && cost256 < cost128 &&
                      (4 * g.BN <= 512 || g.num_k_iters >= 16);
    const bool broad = tiles128_all <= 8L * sms && rounds128 >= 2 && cost256 < cost128 && (4 * g.BN <= 512 || g.num_k_iters >= 16);
    if (area2 > 0 && tiles256 >= 2 && (tune_p2 == 2 || (tune_p2 == 3 && broad) || rule)) {
      g.m2 = 1;
      g.TH = th2;
      g.TW = tw2;
      g.a_stage_bytes = 2 * A_STAGE_BYTES;
    }
  }
  g.tiles_x = ceil_div(OW, g.TW);
  g.tiles_y = ceil_div(OH, g.TH);
  const int m_tiles = B * g.tiles_y * g.tiles_x;
  static const int tune_split = getenv("RDSIC_TC_SPLIT_N") ? atoi(getenv("RDSIC_TC_SPLIT_N")) : 1;
  if (tune_split && groups == 1) {
    // grids below one wave (latent-resolution layers): split N further so that more SMs get a tile.
    // The B tensor map zero-fills rows past ceil16(Cout), so BN need not divide Cout.
    const int c16 = (d->Cout + 15) / 16 * 16;
    if (m_tiles * ceil_div(c16, g.BN) * 2 <= sms) {
      const int n_wanted = sms / m_tiles;  // never more tiles than SMs: a second round costs more than it gains
      int bn = (ceil_div(c16, n_wanted) + 15) / 16 * 16;
      if (bn < 32) bn = 32;
      if (bn < g.BN) g.BN = bn;
    }
  }
  g.n_tiles = ceil_div(d->Cout, g.BN);
  g.total_tiles = m_tiles * g.n_tiles;
  g.m_tiles = m_tiles;
  // MC mode (see TcGeom): CTA pairs share every B stage through TMA multicast.  RDSIC_TC_MC: 0 off, 1 on wherever
  // there are two M tiles (default), 2 only for grids of at least one wave.
  // (measured: correct, and no faster than without -- the shared-memory port, not L2, bounds these tiles -- so off
  // by default; PAIR mode is the one that relieves the port)
  static const int tune_mc = getenv("RDSIC_TC_MC") ? atoi(getenv("RDSIC_TC_MC")) : 0;
  g.pair = want_pair && m_tiles >= 2;
  g.mc = !g.pair && tune_mc && groups == 1 && !g.halo && m_tiles >= 2 && sms >= 2 && (tune_mc != 2 || g.total_tiles >= sms);
  g.walk_total = (g.mc || g.pair) ? ceil_div(m_tiles, 2) * g.n_tiles : g.total_tiles;
  g.b_stage_bytes = (g.pair ? g.BN / 2 : g.BN) * BK * 2;  // PAIR: each CTA stages half of the B box
  const int stage_bytes = g.halo ? g.b_stage_bytes : g.a_stage_bytes + g.b_stage_bytes;
  // one persistent CTA per SM owns the whole shared memory: as deep a TMA ring as fits (the ring keeps
  // running across tiles, so even 2-3-iteration pointwise GEMMs keep many stages in flight)
  static const int tune_stages = getenv("RDSIC_TC_STAGES") ? atoi(getenv("RDSIC_TC_STAGES")) : 0;
  int stages = ((g.m2 ? 220 : 200) * 1024 - (g.halo ? 2 * g.a_halo_bytes : 0)) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (tune_stages > 0 && tune_stages < stages) stages = tune_stages;
  if (!g.halo && (!g.m2 || g.pair)) stages &= ~1;  // even ring depth: stage parity = owner (two producers / two issuers)
  if (stages < 2) return RDSIC_E_ARG;
  g.num_stages = stages;
  // K-split across two issuer warps (see the kernel's header comment) wherever both accumulators fit TMEM.
  // The accumulator set stays double-buffered across tiles while two sets fit the 512 columns (bn_layer <=
  // 128); wider tiles run single-buffered, which only pays when the main loop dwarfs the then exposed
  // epilogue (>= 16 k-iterations).
  // The decision uses only layer properties (Cout, K) -- never the grid-dependent N split above -- so that the
  // fp32 summation order, hence every output bit, is independent of batch size and image size.
  const int bn_layer = groups > 1 ? g.BN : pick_bn(d->Cout);
  g.ksplit = tune_ksplit && !g.m2 && !g.pair && tune_m2 != 3 && !g.halo && g.num_k_iters >= 4 &&
             (4 * bn_layer <= 512 || (tune_ksplit != 3 && 2 * bn_layer <= 512 && g.num_k_iters >= 16));
  g.acc_stride = g.BN * (1 + (g.ksplit | g.m2));
  g.acc_bufs = 2 * g.acc_stride <= 512 ? 2 : 1;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.acc_bufs * g.acc_stride) g.tmem_cols *= 2;
  RDSIC_CHECK_ARG(g.TW * d->stride <= 256 && g.TH * d->stride <= 256);

  // ---- tensor maps
  CUtensorMap ta, tb;
  {
    const cuuint64_t ld_b = (cuuint64_t)d->in.ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)(d->Cin + (groups - 1) * g.a_group_stride), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld_b, ld_b * W, ld_b * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(g.TW * d->stride), (cuuint32_t)(g.TH * d->stride), 1};
    if (g.halo) { box[1] = (cuuint32_t)g.halo_w; box[2] = (cuuint32_t)g.halo_h; }
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    void* base = (void*)((const __nv_bfloat16*)d->in.ptr + d->in.coff);
    CUresult r = encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return RDSIC_E_ARG;
  }
  {
    const int K = d->KH * d->KW * d->Cin;
    const int n_rows = (d->Cout + 15) / 16 * 16;  // the host pads the packed weight to ceil16(Cout) rows
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)((g.mc || g.pair) ? g.BN / 2 : g.BN)};  // MC / PAIR: each CTA of a pair loads half
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d->weight, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return RDSIC_E_ARG;
  }

  const size_t smem = (size_t)stages * stage_bytes + (g.halo ? 2 * g.a_halo_bytes : 0) + 1024 /*align slack*/ +
                      (2 * MAX_STAGES + 8) * 8 + 16 + 16 + (size_t)((d->Cout + 15) / 16 * 16) * 4 /*staged bias*/;
  // vector epilogue: channels-last views whose 16-channel chunks are 32-byte aligned (256-bit LDG/STG)
  auto vec_ok = [](const rdsic_view& v) {
    return !v.ptr || (!v.nchw && v.ld % 16 == 0 && v.coff % 16 == 0 && ((uintptr_t)v.ptr % 32) == 0);
  };
  if (d->pixel_shuffle == 3)  // phase-major pixel shuffle: vector epilogue only
    RDSIC_CHECK_ARG((d->Cout / 4) % 16 == 0 && !d->res.ptr && !d->aux.ptr && !d->out2.ptr && !d->out3.ptr && vec_ok(d->out));
  const bool plain = d->pixel_shuffle != 2 && d->Cout % 16 == 0 && vec_ok(d->out) && vec_ok(d->out2) && vec_ok(d->out3) &&
                     vec_ok(d->res) && vec_ok(d->aux) && (!d->bias || ((uintptr_t)d->bias % 16) == 0);
  ConvTcKernel kern = g.pair ? (plain ? pick_kernel<true, true>(d->epilogue) : pick_kernel<false, true>(d->epilogue))
                             : (plain ? pick_kernel<true, false>(d->epilogue) : pick_kernel<false, false>(d->epilogue));
  // opt in to >48 KB dynamic smem: per (device, kernel); idempotent, so a race between host threads is harmless
  static bool attr_set[16][4][8] = {};
  const int kvar = (plain ? 1 : 0) + (g.pair ? 2 : 0);
  if (!track || !attr_set[dev][kvar][d->epilogue]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    if (track) attr_set[dev][kvar][d->epilogue] = true;
  }
  if (g.mc || g.pair) {
    const int grid = 2 * g.walk_total < (sms & ~1) ? 2 * g.walk_total : (sms & ~1);
    return rdsic_launch(kern, dim3((unsigned)grid), CONV_THREADS, smem, stream, true, ta, tb, dd, g);
  }
  const int grid = g.total_tiles < sms ? g.total_tiles : sms;
  return rdsic_launch(kern, dim3((unsigned)grid), CONV_THREADS, smem, stream, false, ta, tb, dd, g);
}

#ifdef RDSIC_DEBUG
// Profiling aid (not part of the public header): copies the issuer time stamps recorded under RDSIC_TC_DBG_TS.
extern "C" int rdsic_debug_read_ts(long long* host, int n) {
  if (!g_dbg_ts || n > (g_dbg_host ? 16384 : 4096)) return -1;
  if (g_dbg_host) {
    memcpy(host, g_dbg_ts, (size_t)n * sizeof(long long));
    return 0;
  }
  cudaDeviceSynchronize();
  return (int)cudaMemcpy(host, g_dbg_ts, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost);
}
#endif
