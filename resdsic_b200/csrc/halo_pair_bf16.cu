// Narrow stride-1 multi-tap convolution (N <= 16 GEMM columns) on CTA pairs with the input read ONCE: the image head of
// g_s -- deconv(N, 3, k5 s2) as one merged 3x3 convolution with 4 * 3 = 12 columns + PixelShuffle(2) into the NCHW fp32
// x_hat (layers/conv.py: ConvTranspose2d.emit; reference WACNN/utils.py:126-134, cnn.py:128).
//
// Why its own kernel: through conv_tc_kernel this layer streams its input nine times from L2 (im2col by TMA box, one per
// filter tap): 8 GB per launch at batch 24 = 14.5 TB/s of L2->SM traffic for 1 GB of HBM traffic, 510-550 us.  Here, as
// in ru_pair_bf16.cu, a tile's input is one (16+KH-1) x (8+KW-1)-pixel halo patch per 64-channel block (TMA, 128B swizzle)
// whose shifted windows are the taps' A operands, the whole weight (each SM holds half of the 16 rows: 27 KB) is resident,
// and the leader CTA issues cta_group::2 UMMAs (256 x 16 x 16) for both SMs.  L2->SM per 128-row tile: 69 KB instead of
// 432 KB.  MEASURED: 468 us -- and that is the floor of this formulation, not a tuning problem: every UMMA reads its
// 128 x 32-byte A slice from shared memory at ~64 B/cycle/SM (108 MMAs per tile x 64 cycles = 6.9k cycles per tile, the
// measured 7.0k; a second issuer warp changed nothing), the same rate that explains the ~85 cycles per N = 96 MMA of
// the ResidualUnit kernels (4 KB of A + 1.5 KB of B).  With N = 12 the layer is bound by delivering its ACTIVATIONS to the
// tensor core nine times, whatever feeds shared memory.
//
// Roles per CTA (6 warps): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2..5 epilogue (one per TMEM
// lane quarter).  TMEM: two 16-column accumulators (double-buffered across tiles).  (A second issuer warp -- alternate
// taps into a second accumulator -- was measured: no gain, the operand port is the bound; dropped, K runs sequentially.)
// The kernel serves this layer at EVERY map size, so that a pixel's bits do not depend on the batch / image size.
// Barriers: both CTAs' patches complete on the LEADER's a_full[slot]; the leader's commits are multicast to both CTAs'
// a_empty[slot] / acc_full[b]; both CTAs' epilogue warps arrive on the leader's acc_empty[b] (count 8).
#include "tc_common.cuh"

namespace {

constexpr int HP_THREADS = 32 * 6;
constexpr int HP_SLOTS = 6;   // patch ring (u = local tile * kb + channel block; one producer, one issuer, in order)
constexpr int HP_BN = 16;

struct HaloPairGeom {
  int kb, kc_last, kiters, halo_w, halo_h, patch_bytes, patch_tx, b_blk_bytes;
};

__global__ void __launch_bounds__(HP_THREADS, 1)
halo_pair_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const rdsic_conv_desc d,
                    const TcGeom g, const HaloPairGeom hg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* b_res = smem + (size_t)HP_SLOTS * hg.patch_bytes;  // [kiters][8 rows x 128 B]
  uint64_t* a_full = (uint64_t*)(b_res + (size_t)hg.kiters * hg.b_blk_bytes);  // [HP_SLOTS] LEADER's copy
  uint64_t* a_empty = a_full + HP_SLOTS;   // [HP_SLOTS] local, multicast commit
  uint64_t* acc_full = a_empty + HP_SLOTS;  // [2] local, multicast commit
  uint64_t* acc_empty = acc_full + 2;       // [2] LEADER's copy: both CTAs' four epilogue warps
  uint64_t* w_full = acc_empty + 2;         // LEADER's copy: resident weights of both CTAs
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_b) : "memory");
    for (int s = 0; s < HP_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&acc_full[k], 1);
      mbar_init(&acc_empty[k], 8);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  pdl_wait();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();
  const uint32_t tmem_base = *tmem_slot;
  const TileWalk wk = make_walk(g);
  const bool leader = wk.crank == 0;

  if (warp == 0) {
    // ================= TMA producer =================
    const uint32_t patch0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t a_full0 = __shfl_sync(0xffffffffu, smem_u32(a_full), 0);
    const uint32_t a_empty0 = __shfl_sync(0xffffffffu, smem_u32(a_empty), 0);
    const uint32_t lead_a_full0 = mapa_u32(a_full0, 0u);
    int kb = hg.kb, kiters = hg.kiters, total = g.walk_total, step = wk.step, Cin = d.Cin;
    asm volatile("" : "+r"(kb), "+r"(kiters), "+r"(total), "+r"(step), "+r"(Cin));
    {
      const uint32_t wbar = smem_u32(w_full), lead_wbar = mapa_u32(wbar, 0u);
      const uint32_t b0 = __shfl_sync(0xffffffffu, smem_u32(b_res), 0);
      if (elect_one()) {
        if (leader) mbar_expect_tx_u32(wbar, 2u * (uint32_t)(kiters * hg.b_blk_bytes));
        int n = 0;
        for (int tap = 0; tap < d.KH * d.KW; ++tap)
          for (int cb = 0; cb < kb; ++cb, ++n)
            tma_load_2d_2sm_u32(b0 + (uint32_t)(n * hg.b_blk_bytes), &tmap_b, lead_wbar, tap * Cin + cb * BK, wk.crank * (HP_BN / 2));
      }
      __syncwarp();
    }
    uint32_t u = 0;
    for (int q = wk.first; q < total; q += step) {
      int nt, tx, ty, b;
      tile_of(g, wk, q, nt, tx, ty, b);
      const int x0 = tx * g.TW - d.pad_w, y0 = ty * g.TH - d.pad_h;
      for (int cb = 0; cb < kb; ++cb, ++u) {
        const uint32_t slot = u % HP_SLOTS, ph = (u / HP_SLOTS) & 1u;
        mbar_wait_u32(a_empty0 + 8u * slot, ph ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx_u32(a_full0 + 8u * slot, 2u * (uint32_t)hg.patch_tx);
          tma_load_4d_2sm_u32(patch0 + slot * (uint32_t)hg.patch_bytes, &tmap_a, lead_a_full0 + 8u * slot, cb * BK, x0, y0, b);
        }
        __syncwarp();
      }
    }
    // drain: no multicast commit of the leader may arrive on this CTA's barriers after it has exited
    for (int n = 0; n < HP_SLOTS; ++n, ++u) mbar_wait_u32(a_empty0 + 8u * (u % HP_SLOTS), ((u / HP_SLOTS) & 1u) ^ 1u);
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer: leader CTA only =================
    if (leader) {
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a_full0 = __shfl_sync(0xffffffffu, smem_u32(a_full), 0);
      const uint32_t a_empty0 = __shfl_sync(0xffffffffu, smem_u32(a_empty), 0);
      const uint32_t patch_u0 = (__shfl_sync(0xffffffffu, smem_u32(smem), 0) & 0x3FFFFu) >> 4, patch_u = (uint32_t)hg.patch_bytes >> 4;
      const uint32_t b_u0 = (__shfl_sync(0xffffffffu, smem_u32(b_res), 0) & 0x3FFFFu) >> 4, bblk_u = (uint32_t)hg.b_blk_bytes >> 4;
      const uint64_t dconst = make_sw128_desc(0);
      const uint64_t dconst_halo = make_sw128_desc_ex(0, (uint32_t)(hg.halo_w * 128), 0);
      const uint32_t idesc = make_idesc(HP_BN, 2 * BM);
      int kb = hg.kb, total = g.walk_total, step = wk.step, KH = d.KH, KW = d.KW, hw8 = hg.halo_w * 8;
      asm volatile("" : "+r"(kb), "+r"(total), "+r"(step), "+r"(KH), "+r"(KW), "+r"(hw8));
      mbar_wait(w_full, 0);
      tcgen05_fence_after();
      uint32_t lt = 0, u = 0;
      for (int q = wk.first; q < total; q += step, ++lt) {
        const uint32_t b = lt & 1u, use = (lt >> 1) & 1u;
        mbar_wait(&acc_empty[b], use ^ 1u);
        tcgen05_fence_after();
        const uint32_t acc = tbase + b * (uint32_t)HP_BN;
        for (int cb = 0; cb < kb; ++cb, ++u) {
          const uint32_t slot = u % HP_SLOTS, ph = (u / HP_SLOTS) & 1u;
          mbar_wait_u32(a_full0 + 8u * slot, ph);
          tcgen05_fence_after();
          const uint64_t da0 = dconst_halo + (uint64_t)(patch_u0 + slot * patch_u);
          const uint64_t db0 = dconst + (uint64_t)(b_u0 + (uint32_t)cb * bblk_u);
          const int kc = cb + 1 != kb ? 4 : hg.kc_last;
          if (elect_one()) {
            int tap = 0;
            for (int r = 0; r < KH; ++r)
              for (int sx = 0; sx < KW; ++sx, ++tap) {
                const uint64_t da = da0 + (uint64_t)(uint32_t)(r * hw8 + sx * 8);
                const uint64_t db = db0 + (uint64_t)((uint32_t)(tap * kb) * bblk_u);
                for (int k = 0; k < kc; ++k) umma_bf16_2sm(acc, da + 2 * k, db + 2 * k, idesc, (cb | tap | k) ? 1u : 0u);
              }
            tcgen05_commit_2sm_mc_u32(a_empty0 + 8u * slot, 3);
          }
          __syncwarp();
        }
        if (elect_one()) tcgen05_commit_2sm_mc_u32(smem_u32(&acc_full[b]), 3);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: one warp per TMEM lane quarter =================
    const int q = warp % 4;
    const int ml = q * 32 + lane;
    const int dy = ml / g.TW, dx = ml % g.TW;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t lead_acc_empty0 = mapa_u32(smem_u32(acc_empty), 0u);
    const int Cv = d.Cout >> 2;  // output channels after the shuffle
    float* outp = (float*)d.out.ptr;
    uint32_t lt = 0;
    for (int tq = wk.first; tq < g.walk_total; tq += wk.step, ++lt) {
      int nt, tx, ty, b;
      const bool tile_ok = tile_of(g, wk, tq, nt, tx, ty, b);
      const int oy = ty * g.TH + dy, ox = tx * g.TW + dx;
      const bool row_ok = tile_ok && oy < d.OH && ox < d.OW;
      const uint32_t bsel = lt & 1u, par = (lt >> 1) & 1u;
      mbar_wait(&acc_full[bsel], par);
      tcgen05_fence_after();
      float v[16];
      tmem_ld16(tlane + bsel * (uint32_t)HP_BN, v);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_u32(lead_acc_empty0 + 8u * bsel);  // accumulator read: the issuer may reuse it
      if (row_ok) {
        // nn.PixelShuffle(2) into NCHW fp32: column n = 4 c + 2 py + px -> channel c, pixel (2 oy + py, 2 ox + px);
        // the two px of one (c, py) are one 8-byte store, contiguous across the lanes of a patch row
#pragma unroll
        for (int c = 0; c < HP_BN / 4; ++c) {
          if (c >= Cv) break;
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            const int n = 4 * c + 2 * py;
            float2 o;
            o.x = v[n] + (d.bias ? d.bias[n] : 0.f);
            o.y = v[n + 1] + (d.bias ? d.bias[n + 1] : 0.f);
            const size_t idx = (((size_t)b * Cv + c) * d.OHt + (2 * oy + py)) * d.OWt + 2 * ox;
            *reinterpret_cast<float2*>(outp + idx) = o;
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
  }
}

}  // namespace

// Called by rdsic_conv_forward_bf16 after validation; returns -1 when the layer does not qualify.
int rdsic_halo_pair_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream) {
  static const int tune = getenv("RDSIC_HALO_PAIR") ? atoi(getenv("RDSIC_HALO_PAIR")) : 1;
  const int sms = rdsic_sm_count();
  if (!tune || sms < 2 || d->groups > 1 || d->tail_mode || d->epilogue != RDSIC_EPI_NONE) return -1;
  if (d->pixel_shuffle != 2 || d->Cout > HP_BN || d->Cout % 4 || !d->out.nchw || d->out.dtype != RDSIC_F32) return -1;
  if (d->stride != 1 || d->KH > 3 || d->KW > 3 || d->KH * d->KW < 2 || d->Cin % 16 || d->res.ptr || d->aux.ptr || d->out2.ptr ||
      d->out3.ptr || d->OHt != 2 * d->OH || d->OWt != 2 * d->OW || ((uintptr_t)d->out.ptr % 8))
    return -1;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return RDSIC_E_UNSUPPORTED;
  TcGeom g = {};
  const int B = d->B, H = d->H, W = d->W, OH = d->OH, OW = d->OW;
  g.TH = 16;
  g.TW = 8;
  g.tiles_x = ceil_div(OW, g.TW);
  g.tiles_y = ceil_div(OH, g.TH);
  g.BN = HP_BN;
  g.n_tiles = 1;
  g.total_tiles = B * g.tiles_y * g.tiles_x;
  g.m_tiles = g.total_tiles;
  if (g.m_tiles < 2) return -1;
  g.pair = 1;
  g.walk_total = ceil_div(g.m_tiles, 2);
  HaloPairGeom hg;
  hg.kb = ceil_div(d->Cin, BK);
  hg.kiters = d->KH * d->KW * hg.kb;
  hg.kc_last = (d->Cin - (hg.kb - 1) * BK) / 16;
  hg.halo_w = g.TW + d->KW - 1;
  hg.halo_h = g.TH + d->KH - 1;
  hg.patch_tx = hg.halo_w * hg.halo_h * BK * 2;
  hg.patch_bytes = (hg.patch_tx + 1023) / 1024 * 1024;
  hg.b_blk_bytes = (HP_BN / 2) * BK * 2;  // 8 rows x 128 B = one swizzle atom
  const size_t smem = (size_t)HP_SLOTS * hg.patch_bytes + (size_t)hg.kiters * hg.b_blk_bytes + 1024 + (2 * HP_SLOTS + 5) * 8 + 16;
  if (smem > 227u * 1024u) return -1;

  CUtensorMap ta, tb;
  {
    const cuuint64_t ld_b = (cuuint64_t)d->in.ld * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {ld_b, ld_b * W, ld_b * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)hg.halo_w, (cuuint32_t)hg.halo_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    void* base = (void*)((const __nv_bfloat16*)d->in.ptr + d->in.coff);
    if (encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return RDSIC_E_ARG;
  }
  {  // packed weight [16 rows (zero-padded)][K]: each CTA loads its 8 rows of every k-block
    const int K = d->KH * d->KW * d->Cin;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)HP_BN};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(HP_BN / 2)};
    cuuint32_t estr[2] = {1, 1};
    if (encode(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d->weight, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return RDSIC_E_ARG;
  }
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool track = dev >= 0 && dev < 16;
  if (!track || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(halo_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    if (track) attr_set[dev] = true;
  }
  const int grid = 2 * g.walk_total < (sms & ~1) ? 2 * g.walk_total : (sms & ~1);
  return rdsic_launch(halo_pair_tc_kernel, dim3((unsigned)grid), HP_THREADS, smem, stream, true, ta, tb, *d, g, hg);
}
