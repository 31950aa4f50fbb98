// Shared pieces of the tcgen05 / TMA kernels (conv_bf16.cu, conv_gdn_bf16.cu): tile constants, PTX wrappers
// (mbarrier, TMA, tcgen05.mma / ld / st / commit), shared-memory and instruction descriptors, and the
// vectorised epilogue helpers.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 64;         // bf16 elements per 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int NUM_EPI_WARPS = 12;                  // four warps per TMEM lane quarter (latency hiding by TLP)
constexpr int EPI_PARTS = NUM_EPI_WARPS / 4;       // column interleave factor between the warps of a quarter
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;  // TMA warp + MMA warp + epilogue warps
constexpr int MAX_STAGES = 8;

struct TcGeom {
  int TH, TW, tiles_y, tiles_x;
  int BN, n_tiles, total_tiles, kb_per_tap, num_k_iters, num_stages, tmem_cols;
  int b_stage_bytes;
  // halo mode (stride-1 multi-tap convs): the (TH+KH-1) x (TW+KW-1) input patch of a channel block is loaded
  // ONCE and every tap's A operand is a shifted window of it (descriptor start + row offset, SBO = halo row pitch)
  int halo, halo_w, halo_h, a_halo_bytes, halo_base_off;
  int ksplit;          // 1: two issuer warps take alternate k-iterations into two accumulators (summed by the epilogue)
  int acc_bufs;        // accumulator buffers cycled across tiles (2, or 1 when 2 x the accumulator set does not fit TMEM)
  int acc_stride;      // TMEM columns of one buffer = BN * (1 + ksplit)
  int dbg_skip_load;   // profiling aid: bit0 = do not load A, bit1 = do not load B (results are garbage)
  long long* dbg_ts;   // profiling aid: clock64 stamps of issuer 0 of CTA 0 (4 per k-iteration), or null
  // M2 mode (conv_bf16.cu): the tile is a 256-row patch (TH*TW = 256) loaded by ONE TMA box; the two issuer
  // warps each own a 128-row half (own accumulator) and share every B stage, which halves the weight bytes
  // streamed from L2 per output row.  a_stage_bytes = 16 KB (plain) or 32 KB (M2).
  int m2, a_stage_bytes;
  // MC mode (B multicast in CTA pairs): the kernel is launched in clusters of two CTAs that walk pair-tiles -- two
  // M tiles with the SAME n tile -- in lockstep.  Each CTA loads its own A box and HALF of the B box, multicast into
  // the shared memory of both (cp.async.bulk.tensor ... .multicast::cluster), so the weight bytes every SM pulls
  // from L2 halve; every issuer's tcgen05.commit releases the stage in both CTAs (multicast commit), so a stage is
  // refilled only when both SMs have finished reading it.  Pure data movement: the fp32 summation order is unchanged.
  int mc, m_tiles, walk_total;  // walk_total: pair-tiles (mc / pair) or tiles
  // PAIR mode (cta_group::2): the two CTAs of a cluster -- two SMs of one TPC -- execute ONE 256 x BN UMMA per K step.
  // Each CTA stages its own 128-row A box and HALF of the B box (rows [rank * BN/2, +BN/2)); the leader CTA's issuer
  // warp issues tcgen05.mma.cta_group::2 for both, each SM accumulating its 128 rows in its own TMEM.  Per k-iteration
  // an SM's shared-memory port then carries 2 (A + B/2) bytes instead of 2 (A + B) -- the port, not L2, is what bounds
  // the 1-CTA tiles (MC mode halves the L2 bytes of B and measures no faster) -- and one MMA issue covers 256 rows.
  // Both CTAs' TMA loads complete on the LEADER's full barrier; the leader's commits are multicast to both CTAs'
  // empty / accumulator-full barriers; both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.
  int pair;
  // grouped convolution (rdsic_conv_desc.groups): n tile nt is group nt (BN = Cout / groups); its A boxes start at
  // channel nt * a_group_stride of the input view
  int a_group_stride;
};

// walk of a persistent CTA over its tiles: q = first, first + step, ... < g.walk_total; tile_of() maps q to
// (n tile, m tile); m tiles past g.m_tiles (the odd one out of the last pair) decode to an out-of-range image
// index, so their TMA boxes are zero-filled and the epilogue stores nothing
struct TileWalk {
  int first, step, crank;
};
__device__ __forceinline__ TileWalk make_walk(const TcGeom& g) {
  TileWalk w;
  if (g.mc || g.pair) { w.first = (int)(blockIdx.x >> 1); w.step = (int)(gridDim.x >> 1); w.crank = (int)(blockIdx.x & 1u); }
  else { w.first = (int)blockIdx.x; w.step = (int)gridDim.x; w.crank = 0; }
  return w;
}
__device__ __forceinline__ bool tile_of(const TcGeom& g, const TileWalk& w, int q, int& nt, int& tx, int& ty, int& b) {
  nt = q % g.n_tiles;
  int t = q / g.n_tiles;
  if (g.mc || g.pair) t = 2 * t + w.crank;
  const bool valid = t < g.m_tiles;
  tx = t % g.tiles_x;
  t /= g.tiles_x;
  ty = t % g.tiles_y;
  b = t / g.tiles_y;
  return valid;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded mbarrier waits: a lost signal must not hang the GPU box, but a slow wait (profiler replay, debugger,
// co-running kernels) must not kill the context either, so the limit is WALL TIME (%globaltimer, checked every
// 2^14 failed probes), not an iteration count: after MBAR_TIMEOUT_NS the wait gives up with __trap().
constexpr unsigned long long MBAR_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s (a forward step is ~20 ms)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef RDSIC_DEBUG
// Debug builds: host-mapped log for barrier timeouts (one copy per translation unit, set by tc_set_timeout_log).
static __device__ long long* g_mbar_log = nullptr;
static inline void tc_set_timeout_log(long long* p) { cudaMemcpyToSymbol(g_mbar_log, &p, sizeof(p)); }
#endif

__device__ __noinline__ void mbar_timeout(uint32_t addr, uint32_t parity) {
#ifdef RDSIC_DEBUG
  if (g_mbar_log && (threadIdx.x & 31) == 0) {
    volatile long long* rec = g_mbar_log + 4 + ((blockIdx.x % 148) * 16 + (threadIdx.x / 32)) * 4;
    rec[0] = 100;
    rec[1] = ((long long)blockIdx.x << 32) | (threadIdx.x / 32);
    rec[2] = addr;
    rec[3] = parity;
    __threadfence_system();
    __nanosleep(2000000);
  }
#endif
  __trap();
}

__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {  // shared-window address
  uint32_t done = 0, spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 0x3FFFu) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (!t0) t0 = now;
      else if (now - t0 > MBAR_TIMEOUT_NS) mbar_timeout(addr, parity);
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_u32(smem_u32(bar), parity); }
// One lane of a converged warp (elect.sync).  The TMA / MMA loops are run by their WHOLE warp in uniform
// control flow and only the issue itself is predicated: a loop entered by a single lane is divergent code,
// where the compiler cannot use uniform registers and wraps every UTCHMMA / UTMALDG operand in an
// ELECT + R2UR.BROADCAST "waterfall" loop (~150 cycles per MMA issue, measured).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// variants taking shared-window addresses (kept in uniform registers by the callers)
__device__ __forceinline__ void tcgen05_commit_u32(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_u32(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// MC mode: the box goes to the same shared-memory offset of every CTA in `mask`, and completes `bytes` on the
// mbarrier at the same offset of each of them
__device__ __forceinline__ void tma_load_2d_mc_u32(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// arrives (once the issuer's earlier MMAs retire) on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tcgen05_commit_mc_u32(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// ---- cta_group::2 (PAIR mode) forms
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {  // same offset in CTA `cta_rank`'s window
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// Arrive on an mbarrier of (possibly) the peer CTA.  Default semantics (.release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive: what these arrivals hand over is tensor-memory state, ordered by the tcgen05 fences around
// them, never generic memory.  The first version asked for `.release.cluster`: ncu showed the resulting ERRBAR
// (stall_membar) as the hottest line of the pair kernels -- the arriving lane waits until its own STG.256 stores of
// the tile are visible cluster-wide, 1-2k cycles per tile, before it can signal that an accumulator is free.
__device__ __forceinline__ void mbar_arrive_cluster_u32(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in the executing CTA's shared memory, the bytes complete on `bar`, a
// shared::cluster address that may lie in the peer (leader) CTA
__device__ __forceinline__ void tma_load_4d_2sm_u32(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_u32(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm_mc_u32(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// D[tmem, 256 rows over the CTA pair] (+)= A * B^T; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows at 128 B pitch,
// 8-row groups 1024 B apart (SBO), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO [32,46)
  d |= (uint64_t)1 << 46;                       // version = 1 [46,48)
  d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B [61,64)
  return d;
}

// same, with an explicit 8-row-group stride (SBO) and swizzle base offset: used for shifted windows into a
// larger 128B-swizzled tile whose rows are 128 B apart but whose 8-row groups are `sbo_bytes` apart.
__device__ __forceinline__ uint64_t make_sw128_desc_ex(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;          // matrix base offset [49,52)
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=bn
__device__ __forceinline__ uint32_t make_idesc(int bn, int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// tcgen05.ld is asynchronous: issue any number of loads, then tmem_ld_wait(), then pass every destination
// through tmem_ld_fence() -- an empty asm that makes the registers depend on the wait, so the compiler
// cannot hoist their first use above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_fence(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
               "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  tmem_ld16_issue(taddr, r);
  tmem_ld_wait();
  tmem_ld_fence(r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ epilogue helpers
// Epilogue math for the bf16 path.  The pointwise layers are instruction-issue bound in the epilogue (ncu:
// ~34 thread instructions per output element with erff(), profiles/), so GELU is evaluated as
// 0.5x(1 + tanh(x(c1 + c2 x^2 + c3 x^4))) with coefficients fitted to the exact erf form (max |err| 3e-5, 16x
// tighter than the textbook tanh-GELU) on the single-instruction MUFU tanh: < 0.05 bf16 ulp of the stored
// activation and +9 % images/s over erff().  Parity was measured for three variants at 512x768 on the
// low-rate operating point (tests/bpp_experiment.py): per-image bpp error 0.046 % / 0.004 % with this form vs
// 0.059 % / 0.008 % with erff() -- indistinguishable (both are dominated by which symbols flip), so the fast
// form is the default; -DRDSIC_GELU_EXACT (erff/expf) and -DRDSIC_GELU_AS (Abramowitz-Stegun erf) remain
// selectable.  The fp32 mode (conv_f32.cu) always uses erff()/expf().
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
#if defined(RDSIC_GELU_NONE)  // profiling only: what the GELU arithmetic costs
  return x;
#elif defined(RDSIC_GELU_EXACT)
  return gelu_erf(x);
#elif defined(RDSIC_GELU_AS)
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, __expf(-z * z), 1.0f);
  const float h = 0.5f * x;
  return fmaf(h, copysignf(erf_abs, x), h);
#else
  // beyond |x| = 8 the fitted quintic would turn over: freeze x^2 at 64 there, so that t stays 1.70 and tanh(x t)
  // saturates to +-1 (bit-identical to clamping x itself for |x| <= 8, the same +-1 beyond; one FMNMX instead of two)
  const float x2 = fminf(x * x, 64.0f);
  float t = fmaf(-3.58618502e-4f, x2, 3.70495807e-2f);
  t = fmaf(t, x2, 7.97459395e-1f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_mufu(x * t), h);
#endif
}
#if defined(RDSIC_GELU_EXACT) || defined(RDSIC_GELU_AS)
__device__ __forceinline__ float sigmoid_fast(float x) { return sigmoid_f(x); }
#else
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_mufu(0.5f * x), 0.5f); }
#endif

template <int EPI>
__device__ __forceinline__ float epi_apply(float v, float res, float aux) {
  if (EPI == RDSIC_EPI_GELU) return gelu_fast(v);
  if (EPI == RDSIC_EPI_RES_GELU) return gelu_fast(v + res);
  if (EPI == RDSIC_EPI_ADD_RES) return v + res;
  if (EPI == RDSIC_EPI_GATE) return aux * sigmoid_fast(v) + res;
  if (EPI == RDSIC_EPI_GDN) return res * rsqrtf(v);
  if (EPI == RDSIC_EPI_IGDN) return res * sqrtf(v);
  if (EPI == RDSIC_EPI_LRP) return res + 0.5f * tanhf(v);
  return v;
}

// explicit shared-window 128-bit load (bias vectors staged per CTA): through a generic pointer the compiler emits LD.E.128,
// whose longer round trip showed as long-scoreboard stalls on the first FADD of every epilogue chunk (ncu, round 2)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 f;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(addr));
  return f;
}

// 16 consecutive channels of one pixel = one 32-byte (bf16) or two 32-byte (fp32) accesses; sm_100 has
// 256-bit LDG/STG, and the host guarantees 32-byte alignment of every chunk on the PLAIN path.
struct Pack8 {
  uint32_t w[8];
};
__device__ __forceinline__ Pack8 ldg256(const void* p) {
  Pack8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(void* p, const Pack8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}
__device__ __forceinline__ void unpack_bf16x16(const Pack8& r, float* o) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[2 * i] = __uint_as_float(r.w[i] << 16);
    o[2 * i + 1] = __uint_as_float(r.w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void load16(const rdsic_view& vw, size_t elem, float* o) {
  if (vw.dtype == RDSIC_BF16) {
    unpack_bf16x16(ldg256((const __nv_bfloat16*)vw.ptr + elem), o);
  } else {
    const Pack8 r0 = ldg256((const float*)vw.ptr + elem), r1 = ldg256((const float*)vw.ptr + elem + 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = __uint_as_float(r0.w[i]);
      o[8 + i] = __uint_as_float(r1.w[i]);
    }
  }
}
__device__ __forceinline__ void store16(const rdsic_view& vw, size_t elem, const float* v, bool sq) {
  if (vw.dtype == RDSIC_BF16) {
    Pack8 r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a0 = v[2 * i], a1 = v[2 * i + 1];
      if (sq) { a0 *= a0; a1 *= a1; }
      __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
      r.w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    stg256((__nv_bfloat16*)vw.ptr + elem, r);
  } else {
    Pack8 r0, r1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r0.w[i] = __float_as_uint(sq ? v[i] * v[i] : v[i]);
      r1.w[i] = __float_as_uint(sq ? v[8 + i] * v[8 + i] : v[8 + i]);
    }
    stg256((float*)vw.ptr + elem, r0);
    stg256((float*)vw.ptr + elem + 8, r1);
  }
}


// ------------------------------------------------------------------ host: TMA descriptor encoder
// resolved through the runtime (no link-time dependency on libcuda: the library loads without a driver)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;  // benign race: every thread resolves the same pointer
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}


}  // namespace
