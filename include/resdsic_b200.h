/* resdsic_b200 -- C ABI of the B200-native WACNN/STF codec forward path.
 *
 * The reference (AlbertoPresta/ResDSIC) has NO FFI on this path: its boundary
 * is the Python nn.Module surface.  This header is what a
 * maintainer would bind (ctypes) to replace the ATen calls
 * behind each reference module; every entry point cites the reference code it
 * replaces (paths relative to the reference's src/compress/).
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch types;
 *   - every function is re-entrant, launches on the caller's `stream` of the
 *     CURRENT device and never allocates or retains caller memory (outputs /
 *     workspaces are caller-owned).  The only process-wide state is a set of
 *     idempotent caches written once with the same value by whichever host
 *     thread gets there first: per-device SM count, "dynamic shared memory
 *     opt-in done" flags per kernel, the resolved driver entry point of the
 *     tensor-map encoder, and developer tuning knobs (RDSIC_TC_* environment
 *     variables) read once.  Profiling / tracing hooks exist only in builds
 *     compiled with -DRDSIC_DEBUG;
 *   - return value: 0 = ok, otherwise a negative RDSIC_E_* argument error or a
 *     positive cudaError_t; rdsic_error_string() renders either.  The Python
 *     host raises RuntimeError on non-zero, matching the reference's
 *     exception-based error convention (entropy_models.py:129-130,184-203);
 *   - activations are channels-last (NHWC) with an explicit per-pixel channel
 *     stride `ld` and channel offset `coff`, so torch.cat / chunk / PixelShuffle /
 *     roll / window_partition are folded into addressing, never materialised.
 */
#ifndef RESDSIC_B200_H
#define RESDSIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDSIC_ABI_VERSION 8

typedef void* rdsic_stream_t; /* cudaStream_t */

enum { RDSIC_F32 = 0, RDSIC_BF16 = 1 };

enum {
  RDSIC_E_ARG = -1,      /* invalid argument / unsupported shape */
  RDSIC_E_ALIGN = -2,    /* pointer or stride not aligned as required */
  RDSIC_E_UNSUPPORTED = -3
};

/* Epilogues fused into the convolution / GEMM kernels. v = acc + bias. */
enum {
  RDSIC_EPI_NONE = 0,     /* out = v                                  nn.Conv2d */
  RDSIC_EPI_GELU = 1,     /* out = gelu(v)            conv + nn.GELU (layers.py:58-63, cnn.py:56-129) */
  RDSIC_EPI_RES_GELU = 2, /* out = gelu(v + res)      ResidualUnit tail (layers.py:67-71) */
  RDSIC_EPI_ADD_RES = 3,  /* out = v + res            attention proj + shortcut (win_attention.py:113,205) */
  RDSIC_EPI_GATE = 4,     /* out = aux*sigmoid(v)+res Win_noShift_Attention.forward (layers.py:83-89) */
  RDSIC_EPI_GDN = 5,      /* out = res*rsqrt(v)       GDN.forward (gdn.py:62-75), A operand squared */
  RDSIC_EPI_IGDN = 6,     /* out = res*sqrt(v)        inverse GDN (gdn.py:70-71) */
  RDSIC_EPI_LRP = 7       /* out = res + 0.5*tanh(v)  latent residual prediction (cnn.py:179-182) */
};

/* One tensor view: NHWC element (b,y,x,c) lives at
 * ptr[((b*H + y)*W + x)*ld + coff + c]; with nchw != 0 it is
 * ptr[((b*C + c)*H + y)*W + x] (ld/coff ignored, C = channel count of the op) -- except in
 * rdsic_copy_forward, where an NCHW view may address a channel range of a wider tensor:
 * ptr[((b*ld + coff + c)*H + y)*W + x] with ld = channel count of the underlying tensor (0 = C). */
typedef struct rdsic_view {
  void* ptr;
  int32_t dtype; /* RDSIC_F32 | RDSIC_BF16 */
  int32_t ld;
  int32_t coff;
  int32_t nchw;
} rdsic_view;

/* Implicit-GEMM convolution: replaces nn.Conv2d / nn.ConvTranspose2d(phase) /
 * nn.Linear / GDN's 1x1 contraction + the elementwise ops listed above.
 *   M = B*OH*OW output positions, N = Cout, K = KH*KW*Cin.
 *   input pixel for tap (r,s): (oy*stride - pad_h + r, ox*stride - pad_w + s), zero outside.
 *   output pixel: (oy*osy + ooy, ox*osx + oox) of an [B,OHt,OWt] tensor; with
 *   pixel_shuffle = 2, GEMM column n goes to channel n/4 of pixel
 *   (2*oy + (n%4)/2, 2*ox + n%2)  (nn.PixelShuffle, layers.py:34-38).
 * A stride-2 5x5 transposed convolution (WACNN/utils.py:126-134) is four such
 * descriptors (phase sub-kernels 3x3, 3x2, 2x3, 2x2 with osy = osx = 2). */
typedef struct rdsic_conv_desc {
  rdsic_view in;          /* [B,H,W,Cin] */
  int32_t B, H, W, Cin;
  const void* weight;     /* packed [Cout][KH*KW*Cin] (tap-major, channel-minor), compute dtype */
  const float* bias;      /* [Cout] fp32 or NULL */
  int32_t w_dtype;
  int32_t Cout, KH, KW, stride, pad_h, pad_w;
  int32_t OH, OW;         /* GEMM output grid */
  int32_t OHt, OWt, osy, osx, ooy, oox;
  int32_t pixel_shuffle;  /* 0; 2: nn.PixelShuffle(2), GEMM column n = 4 c + s; 3: the same with the weight rows packed
                           * sub-position-major, column n = s * Cout/4 + c (bf16 path: 16 consecutive columns are then
                           * 16 consecutive channels of ONE output pixel, i.e. one 32-byte store) */
  int32_t epilogue;       /* RDSIC_EPI_* */
  int32_t a_square;       /* square the A operand (GDN) */
  rdsic_view out;         /* [B,OHt,OWt,Cout(/4)] */
  rdsic_view res;         /* residual / multiplicand, same pixel grid + channels as out */
  rdsic_view aux;         /* gate operand `a` */
  rdsic_view out2;        /* optional extra copies of the result (slice-loop support buffers) */
  rdsic_view out3;
  int32_t out2_square;    /* out2 receives result^2 (feeds the next GDN's beta + gamma @ x^2 contraction) */
  /* Fused second ("tail") GEMM, bf16 tensor-core path only: the conv result x = conv(in) + bias never leaves
   * the SM; a pointwise GEMM with `tail_weight` (packed bf16 [tail_n][Cout]) runs on it from tensor memory:
   *   tail_mode 1 / 2: GDN / inverse GDN, out = x * rsqrt|sqrt(tail_bias + tail_weight @ x^2)
   *                    (layers/gdn.py:62-75 after WACNN/utils.py:116-134; tail_weight = gamma', tail_n = Cout);
   *   tail_mode 3:     ResidualUnit tail, out = gelu(tail_weight @ gelu(x) + tail_bias + res)
   *                    (layers/layers.py:58-71: conv3x3 -> GELU -> conv1x1, += identity, GELU).
   * `out` (and `res`) then have tail_n channels. */
  int32_t tail_mode;
  const void* tail_weight;
  const float* tail_bias;
  int32_t tail_n;
  /* Grouped form (bf16 tensor-core path only; groups <= 1: an ordinary convolution).  The Cout output channels are
   * `groups` equal blocks; block g is an independent convolution with ITS OWN Cin input channels, which start at channel
   * in.coff + g * in_group_stride of the input view (0: every group reads the same channels), and its own rows of the
   * packed weight [Cout][KH*KW*Cin] / bias / res / out channels.  One launch then replaces `groups` launches of identical
   * shape -- the cc_mean_i / cc_scale_i pairs and the mutually independent slices 5..9 of the slice loop
   * (cnn.py:165-184): per-launch floor ~11 us against ~1 us of tensor time for the narrow layers. */
  int32_t groups;
  int32_t in_group_stride;
  int32_t pad_;
} rdsic_conv_desc;

/* Fused shifted-window attention core: replaces roll + window_partition +
 * (q*scale)@k^T + relative-position bias + shift mask(-100) + softmax + @v +
 * window_reverse + roll back  (win_attention.py:94-112,159-200).
 * qkv is the unshifted [B,H,W,3C] output of the qkv 1x1 GEMM (channel =
 * which*C + head*d + k, win_attention.py:91-92); out is [B,H,W,C] at the
 * ORIGINAL pixel positions, channel = head*d + k. */
typedef struct rdsic_attn_desc {
  rdsic_view qkv;
  rdsic_view out;
  const float* bias_table; /* [(2*ws-1)^2, heads] fp32 (win_attention.py:60-61) */
  int32_t B, H, W, C, heads, ws, shift;
  float scale;             /* qk_scale or head_dim**-0.5 (win_attention.py:54), applied to q before q@k^T */
} rdsic_attn_desc;

/* EntropyBottleneck.forward (entropy_models.py:447-490) + the ste_round z_hat of
 * cnn.py:152-154 (same value in eval mode).  params: [C][RDSIC_EB_STRIDE]
 * fp32 per channel = softplus(matrix0..4) (3,9,9,9,3) | bias0..4 (3,3,3,3,1) |
 * tanh(factor0..3) (3 each) | median.
 * Training mode (quantize "noise", entropy_models.py:131-137): with noise.ptr != NULL the
 * likelihood is evaluated at z + noise instead of round(z-med)+med; z_hat (and symbols)
 * keep the ste_round value, which is what WACNN.forward feeds the hyper-synthesis in
 * both modes; noisy_out (optional) receives z + noise, the module's first output. */
#define RDSIC_EB_STRIDE 60
typedef struct rdsic_eb_desc {
  rdsic_view z;       /* [B,h,w,C] fp32 */
  rdsic_view z_hat;   /* [B,h,w,C] */
  float* lik;         /* [B,C,h,w] fp32 NCHW (module output) */
  int32_t* symbols;   /* optional [B,C,h,w] int32: round(z - median) */
  const float* params;
  int32_t B, h, w, C;
  float lik_bound;    /* 1e-9 */
  int32_t pad_;
  rdsic_view noise;     /* optional [B,h,w,C] fp32, U(-1/2,1/2) drawn by the caller */
  rdsic_view noisy_out; /* optional [B,h,w,C] fp32: z + noise */
} rdsic_eb_desc;

/* GaussianConditional.forward (eval) + ste_round + build_indexes + quantize
 * "symbols" for one channel slice (entropy_models.py:627-668,139-152;
 * cnn.py:175-177,253-254).  Per element:
 *   y_hat = rint(y-mu)+mu; lik = max(.5erfc(c(.5-|y_hat-mu|)/s) - .5erfc(c(-.5-|..|)/s), bound),
 *   s = max(scale, scale_bound); sym = int(rint(y-mu)); idx = #{t in table[:-1] : t < s}. */
typedef struct rdsic_gc_desc {
  rdsic_view y;        /* [B,h,w,Cs] fp32 slice view */
  rdsic_view mu;       /* fp32 */
  rdsic_view scale;    /* fp32 */
  rdsic_view y_hat[3]; /* up to three destinations (ptr NULL = unused) */
  float* lik;          /* NCHW [B,Ctot,h,w] base; slice starts at channel lik_coff */
  int32_t* symbols;    /* optional, same layout */
  int32_t* indexes;    /* optional, same layout */
  const float* table;  /* [n_table] fp32 scale table (cnn.py:14-20) */
  int32_t n_table;
  int32_t B, h, w, Cs, Ctot, lik_coff;
  float scale_bound, lik_bound;
  /* ResDSIC `cimd` (scalable/conditional_multiple_decoder.py:210): when non-zero the conditional is evaluated with
   * scale * mask + scale_eps (the reference adds 1e-7 to the masked scale before the likelihood); 0 = off. */
  float scale_eps;
  /* Training mode (GaussianConditional.forward with quantize "noise", entropy_models.py:131-137,
   * 646-661): with noise.ptr != NULL, lik is evaluated at |y + noise - mu|; y_hat / symbols keep
   * the ste_round value (cnn.py:177); noisy_out (optional) receives y + noise. */
  rdsic_view noise;     /* optional [B,h,w,Cs] fp32 */
  rdsic_view noisy_out; /* optional [B,h,w,Cs] fp32 */
  /* Decoder side (GaussianConditional.dequantize after the rANS decode, cnn.py:326-328, entropy_models.py:160-167):
   * with sym_in != NULL the quantised residual is READ from this int32 tensor (same NCHW [B,Ctot,h,w] layout and
   * channel offset as `symbols`) instead of being computed from y, i.e. y_hat = float(sym_in) + mu; y is ignored
   * (may be NULL).  lik then is the likelihood of the decoded symbol. */
  const int32_t* sym_in;
  /* ResDSIC progressive stream (scalable/single_decoder.py:447-453): with mask.ptr != NULL (fp32 view like y,
   * values 0/1 after Mask.apply_noise's round, layers/mask_layer.py:32-39) the conditional is evaluated with
   * scale * mask, the stored y_hat is rint(y - mu) * mask + mu and symbols = int(rint(y - mu) * mask); the
   * likelihood is still taken at rint(y - mu) + mu, as `gaussian_conditional_prog(y, scale * mask, mu)` does. */
  rdsic_view mask;
} rdsic_gc_desc;

/* ResDSIC importance mask (layers/mask_layer.py:41-107, eval mode: Mask.forward followed by apply_noise's round).
 *   mode 1 "learnable-mask-gamma":  mask = rint(pow(sigmoid(in[0]), gamma[c]))        (:64-90; gamma[c] =
 *          relu(sum of the first scalable_levels-1-pr rows of Mask.gamma) + 1e-7, computed by the caller)
 *   mode 2 "learnable-mask-nested": mask = rint(sigmoid(sum_i sigmoid(in[i])))         (:92-107, n_in = pr)
 * in[i] are the fp32 outputs of Mask.mask_conv (1x1 convolutions over cat(scale, scale_prog)). */
typedef struct rdsic_mask_desc {
  rdsic_view in[8];
  rdsic_view out;      /* fp32 [B,H,W,C] */
  const float* gamma;  /* [C] fp32 (mode 1) */
  int32_t n_in, mode;
  int32_t B, H, W, C;
} rdsic_mask_desc;

/* Layout / elementwise helpers used by the standalone module API. */
typedef struct rdsic_copy_desc {
  rdsic_view src;
  rdsic_view dst;
  int32_t B, H, W, C;
  int32_t op; /* 0 copy/cast, 1 gelu, 2 square, 3 clamp to [0,1] (decompress: x_hat.clamp_(0, 1), cnn.py:340),
                 4 dst = src + src2 (ResDSIC: y_hat_complete = y_hat + y_hat_prog, scalable/single_decoder.py:472) */
  int32_t pad_;
  rdsic_view src2; /* op 4 only; same geometry as src */
} rdsic_copy_desc;

/* LayerNorm over channels (stf Swin block, TCM/tcm.py:214-236). */
typedef struct rdsic_ln_desc {
  rdsic_view in;
  rdsic_view out;
  const float* gamma;
  const float* beta;
  int32_t rows, C;
  float eps;
} rdsic_ln_desc;

/* im2col of a narrow-channel input (the 3-channel image, cnn.py:32) into a channels-last bf16 patch
 * tensor [B,OH,OW,Kp]: element k = (r*KW + s)*C + c for k < KH*KW*C, zero padding up to Kp (multiple
 * of 16).  The first 5x5 s2 convolution then runs as a pointwise tensor-core GEMM with K = Kp. */
typedef struct rdsic_patch_desc {
  rdsic_view src;  /* [B,C,H,W] NCHW or NHWC, fp32 or bf16 */
  rdsic_view dst;  /* [B,OH,OW,Kp] bf16 or fp32 */
  int32_t B, H, W, C, KH, KW, stride, pad, OH, OW, Kp;
  int32_t pad_;
} rdsic_patch_desc;

enum { RDSIC_OP_CONV = 0, RDSIC_OP_ATTN = 1, RDSIC_OP_EB = 2, RDSIC_OP_GC = 3, RDSIC_OP_COPY = 4, RDSIC_OP_LN = 5, RDSIC_OP_PATCH = 6,
       RDSIC_OP_FORK = 7, RDSIC_OP_JOIN = 8, RDSIC_OP_RECORD = 9, RDSIC_OP_WAIT = 10, RDSIC_OP_MASK = 11 };

/* One node of a forward "program" (WACNN.forward, cnn.py:143-193, is ~330 of these). */
/* Independent branches (cc_mean || cc_scale, h_mean_s || h_scale_s, conv_a || conv_b, the context stacks of
 * slices 5..9) may run concurrently on up to RDSIC_MAX_LANES lanes.  A SYNC op (kinds FORK / JOIN are the
 * same operation, named for readability) makes lane `op.lane` wait for everything issued so far on lane
 * `op.u.sync.src`; every other op runs on its `lane`.  rdsic_run_program() executes all lanes in program
 * order on the one stream it is given (always a valid schedule); the CUDA-graph form turns lanes into
 * parallel graph branches and joins every lane back into lane 0 at the end.
 * RECORD marks the current position of lane `op.lane` as event `u.sync.event`; a later WAIT makes its lane
 * wait for exactly that position (decoupled fork/join, used to consume work pre-computed on another lane). */
#define RDSIC_MAX_LANES 12
#define RDSIC_MAX_EVENTS 128
typedef struct rdsic_sync_desc {
  int32_t src;   /* FORK / JOIN: lane to wait for */
  int32_t event; /* RECORD / WAIT: event id in [0, RDSIC_MAX_EVENTS) */
} rdsic_sync_desc;

typedef struct rdsic_op {
  int32_t kind;
  int32_t lane;
  union {
    rdsic_conv_desc conv;
    rdsic_attn_desc attn;
    rdsic_eb_desc eb;
    rdsic_gc_desc gc;
    rdsic_copy_desc copy;
    rdsic_ln_desc ln;
    rdsic_patch_desc patch;
    rdsic_sync_desc sync;
    rdsic_mask_desc mask;
  } u;
} rdsic_op;

int rdsic_abi_version(void);
const char* rdsic_error_string(int code);
/* sizeof(rdsic_op) etc., so the host binding can verify its struct mirror. */
int rdsic_sizeof(int what); /* 0 op, 1 conv, 2 attn, 3 eb, 4 gc, 5 copy, 6 view, 7 ln, 8 patch, 9 mask */

int rdsic_conv_forward(const rdsic_conv_desc* d, rdsic_stream_t stream);
int rdsic_attn_forward(const rdsic_attn_desc* d, rdsic_stream_t stream);
int rdsic_eb_forward(const rdsic_eb_desc* d, rdsic_stream_t stream);
int rdsic_gc_forward(const rdsic_gc_desc* d, rdsic_stream_t stream);
/* EntropyBottleneck.loss (entropy_models.py:396-399), summed by CompressionModel.aux_loss
 * (WACNN/base.py:22-27): terms[c*3+k] = |logits_cumulative(quantiles[c][k]) - target[k]| and
 * *sum = their total (fixed-order tree sum, deterministic).  params as in rdsic_eb_desc;
 * quantiles [C][3], target [3], terms [C*3] and sum [1] are fp32 device pointers. */
int rdsic_eb_aux_loss(const float* params, const float* quantiles, const float* target, int32_t C, float* terms,
                      float* sum, rdsic_stream_t stream);

/* ---- CDF tables for the entropy coder: `update()`.  The tables are what the reference's rANS
 * coder consumes (`_quantized_cdf` int32 [rows, max_length+2], `_cdf_length`, `_offset`); building them is a
 * once-per-model step.  Three stages, all on the device:
 *   sizes  per row: support of the pmf  -> offset[row], cdf_length[row] (= pmf_length + 2)
 *   pmf    prob[row][0..pmf_length) = pmf, prob[row][pmf_length] = tail mass (float32, reference op order)
 *   cdf    pmf_to_quantized_cdf (pip compressai `_CXX`, C++ absent from the reference tree: restated from the
 *          published algorithm, integer-exact against oracle/cdf_oracle.py) applied to every row.          */
/* GaussianConditional.update (entropy_models.py:599-625): center = ceil(table*multiplier),
 * multiplier = -norm.ppf(tail_mass/2) computed by the caller; offset = -center, cdf_length = 2*center+3. */
int rdsic_gc_cdf_sizes(const float* table, int32_t rows, float multiplier, int32_t* offset, int32_t* cdf_length,
                       rdsic_stream_t stream);
int rdsic_gc_pmf(const float* table, const int32_t* offset, int32_t rows, float* prob, int32_t ld, rdsic_stream_t stream);
/* EntropyBottleneck.update (entropy_models.py:356-394): quantiles [C][3]; minima/maxima = clamp(ceil(.),0);
 * offset = -minima, cdf_length = maxima+minima+3.  params as in rdsic_eb_desc; max_length = max pmf_length
 * (the tail mass uses the sample at max_length-1, entropy_models.py:388). */
int rdsic_eb_cdf_sizes(const float* quantiles, int32_t C, int32_t* offset, int32_t* cdf_length, rdsic_stream_t stream);
int rdsic_eb_pmf(const float* params, const float* quantiles, const int32_t* offset, const int32_t* cdf_length, int32_t C,
                 int32_t max_length, float* prob, int32_t ld, rdsic_stream_t stream);
/* EntropyModel._pmf_to_cdf (entropy_models.py:174-182) + pmf_to_quantized_cdf: row r uses prob[r*ld .. +cdf_length[r]-1)
 * and writes cdf[r*cdf_ld .. +cdf_length[r]), zero padded to cdf_ld.  *status (device int32, caller-zeroed) is set
 * to 1 + row if a row is invalid (negative / non-finite / all-zero pmf, or no symbol left to steal from). */
int rdsic_pmf_to_quantized_cdf(const float* prob, int32_t ld, const int32_t* cdf_length, int32_t rows, int32_t precision,
                               int32_t* cdf, int32_t cdf_ld, int32_t* status, rdsic_stream_t stream);

int rdsic_copy_forward(const rdsic_copy_desc* d, rdsic_stream_t stream);
int rdsic_ln_forward(const rdsic_ln_desc* d, rdsic_stream_t stream);
int rdsic_patch_forward(const rdsic_patch_desc* d, rdsic_stream_t stream);
int rdsic_mask_forward(const rdsic_mask_desc* d, rdsic_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training: backward ("_bwd") twins of the forward kernels, fp32 (BASELINE config 4; SURVEY 8b "+ _bwd twins").
 * They replace what torch autograd runs for the reference's training step (training/step.py:42-56 with
 * RateDistortionLoss, training/loss.py:6-30); resdsic_b200/training/functions.py wraps each forward / backward
 * pair as a torch.autograd.Function.  All tensors fp32; `g_*` / `d*` are gradients. */

/* d input of a convolution: in = d x [B,H,W,Cin] (WRITTEN), out = d(conv output) [B,OH,OW,Cout] (read, Cout % 16 == 0),
 * geometry (KH, KW, stride, pad, OH, OW) of the FORWARD conv; wt_dgrad = the forward weight packed [Cin][KH*KW*Cout]
 * (tap-major, output-channel-minor).  Also the backward of nn.Linear / the GDN contraction (1x1), and -- with the
 * roles of in/out swapped by the caller -- nothing else: a transposed conv's d input is rdsic_conv_forward. */
int rdsic_conv_dgrad_f32(const rdsic_conv_desc* d, const float* wt_dgrad, rdsic_stream_t stream);
/* d weight (+ d bias, optional): in = x (read, NHWC or NCHW), out = d(conv output) (read); dw [Cout][KH*KW*Cin] in the
 * forward's packed layout and db [Cout] are zeroed, then accumulated with fp32 atomics (pixels split over CTAs).
 * A transposed conv's d weight is the same call with x and d out exchanged (WACNN/utils.py:126-134). */
int rdsic_conv_wgrad_f32(const rdsic_conv_desc* d, float* dw, float* db, rdsic_stream_t stream);

/* Elementwise forward / backward nodes over flat fp32 arrays of n elements (unused operands NULL):
 *   ADD o0=a+b | GELU_FWD o0=gelu(a) | GELU_BWD o0=a*gelu'(b) | GATE_FWD o0=a*sigmoid(b)+c (layers.py:83-89) |
 *   GATE_BWD (a=g, b=gate operand, c=logits) o0=d operand, o1=d logits | GDN_FWD o0=a*rsqrt(b) (alpha>0: a*sqrt(b);
 *   gdn.py:70-75) | GDN_BWD (a=g, b=x, c=norm) o0=d x (direct term), o1=d norm | SQUARE_FWD o0=a^2 | SQUARE_BWD
 *   o0=2*a*b | LRP_FWD o0=a+0.5*tanh(b) (cnn.py:179-182) | LRP_BWD o0=a*0.5*(1-tanh(b)^2) | RECIP_SCALE o0=alpha/a |
 *   DIFF_SCALE o0=alpha*(a-b) | SCALE o0=alpha*a | MUL o0=a*b. */
enum { RDSIC_PW_ADD = 0, RDSIC_PW_GELU_FWD, RDSIC_PW_GELU_BWD, RDSIC_PW_GATE_FWD, RDSIC_PW_GATE_BWD, RDSIC_PW_GDN_FWD,
       RDSIC_PW_GDN_BWD, RDSIC_PW_SQUARE_FWD, RDSIC_PW_SQUARE_BWD, RDSIC_PW_LRP_FWD, RDSIC_PW_LRP_BWD,
       RDSIC_PW_RECIP_SCALE, RDSIC_PW_DIFF_SCALE, RDSIC_PW_SCALE, RDSIC_PW_MUL };
int rdsic_pointwise_f32(int op, size_t n, const float* a, const float* b, const float* c, float* o0, float* o1, float alpha,
                        rdsic_stream_t stream);
/* nn.PixelShuffle(2) on NHWC: in [B,H,W,4C] -> out [B,2H,2W,C]; inverse != 0: the other way (its backward). */
int rdsic_pixel_shuffle_f32(int inverse, const float* in, float* out, int B, int H, int W, int C, rdsic_stream_t stream);

/* Backward of rdsic_attn_forward (win_attention.py:94-112): d->qkv = the forward's qkv, dout = d out [B,H,W,C];
 * writes dqkv [B,H,W,3C] (every element once) and accumulates d relative_position_bias_table
 * [(2*ws-1)^2][heads] (zeroed first). */
int rdsic_attn_backward_f32(const rdsic_attn_desc* d, const rdsic_view* dout, const rdsic_view* dqkv, float* dbias,
                            rdsic_stream_t stream);

/* Backward of rdsic_gc_forward in noise (training) mode: d carries the forward's y / mu / scale / noise views and
 * bounds; g_lik has the layout of d->lik (NCHW [B,Ctot,h,w], slice at lik_coff); g_yhat (optional) = gradient of
 * y_hat = ste_round(y - mu) + mu (cnn.py:177; identity to y, zero to mu).  LowerBound rule of ops/bound_ops.py:25-27
 * on scale (0.11) and on the likelihood (1e-9).  Writes dy, dmu, dscale [B,h,w,Cs]. */
int rdsic_gc_backward(const rdsic_gc_desc* d, const float* g_lik, const rdsic_view* g_yhat, const rdsic_view* dy,
                      const rdsic_view* dmu, const rdsic_view* dscale, rdsic_stream_t stream);
/* Backward of rdsic_eb_forward in noise mode (entropy_models.py:401-434, sign detached :429-430): g_lik NCHW [B,C,h,w];
 * g_zhat (optional) = gradient of z_hat = ste_round(z - median) + median (identity to z); writes dz [B,h,w,C] and
 * dparams [C][RDSIC_EB_STRIDE] = gradient w.r.t. the PACKED parameters (softplus(matrix) | bias | tanh(factor)). */
int rdsic_eb_backward(const rdsic_eb_desc* d, const float* g_lik, const rdsic_view* g_zhat, const rdsic_view* dz,
                      float* dparams, rdsic_stream_t stream);
/* Backward of rdsic_eb_aux_loss w.r.t. `quantiles` only (entropy_models.py:396-399): dquantiles [C][3], g = d loss. */
int rdsic_eb_aux_backward(const float* params, const float* quantiles, const float* target, int C, float g,
                          float* dquantiles, rdsic_stream_t stream);
/* Reductions of RateDistortionLoss (training/loss.py:24-28), fp64 accumulation into *out (device, zeroed here):
 *   SUM sum(a) | SUM_LOG sum(log a) | SSE sum((a-b)^2). */
enum { RDSIC_RED_SUM = 0, RDSIC_RED_SUM_LOG = 1, RDSIC_RED_SSE = 2 };
int rdsic_reduce_f32(int op, size_t n, const float* a, const float* b, double* out, rdsic_stream_t stream);

/*
 * Host-facing image I/O (csrc/image_io.cu): what the reference's evaluation loop (eval_model/__main__.py:133-147) does
 * on the host with fp32 images and full likelihood tensors, done on the device so that 1 byte per sample and 8 bytes
 * per image cross PCIe.  All pointers are device pointers, 16-byte aligned; NCHW contiguous tensors of n samples.
 *   u8_to_f32: dst = src / 255 (torchvision ToTensor)      f32_to_u8: dst = round(clamp(src, 0, 1) * 255)
 *   rate_per_image: bits[b] = -(sum log2 lik_y[b] + sum log2 lik_z[b]) -- the rate term of RateDistortionLoss
 *   (training/loss.py:14-22) per image before the division by the pixel count; ny / nz = elements per image;
 *   workspace: rdsic_rate_workspace_doubles(B) doubles; fp64, fixed summation order.
 */
int rdsic_image_u8_to_f32(const uint8_t* src, float* dst, size_t n, rdsic_stream_t stream);
int rdsic_image_f32_to_u8(const float* src, uint8_t* dst, size_t n, rdsic_stream_t stream);
int rdsic_rate_per_image(const float* lik_y, size_t ny, const float* lik_z, size_t nz, int B, double* workspace, double* bits,
                         rdsic_stream_t stream);
int rdsic_rate_workspace_doubles(int B);

/* Launch a whole program in order on `stream`.  *n_launched (optional) receives
 * the number of kernels launched.  Stops at the first error; *failed_op
 * (optional) receives its index. */
int rdsic_run_program(const rdsic_op* ops, int n_ops, rdsic_stream_t stream, int* n_launched, int* failed_op);

/* Capture a program into a CUDA graph once and replay it (pointers are baked in:
 * the host keeps its buffers static).  The handle is owned by the caller. */
typedef struct rdsic_graph rdsic_graph;
int rdsic_graph_create(const rdsic_op* ops, int n_ops, rdsic_stream_t stream, rdsic_graph** out);
int rdsic_graph_launch(rdsic_graph* g, rdsic_stream_t stream);
int rdsic_graph_num_kernels(const rdsic_graph* g);
void rdsic_graph_destroy(rdsic_graph* g);

#ifdef __cplusplus
}
#endif
#endif /* RESDSIC_B200_H */
