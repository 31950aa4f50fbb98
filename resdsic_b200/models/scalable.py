"""ResDSIC scalable codecs: the reference's `-m icd` / `-m imd` / `-m cicd` / `-m cimd` / `-m ind`
models (models/WACNN/scalable/single_decoder.py:25-504, multiple_decoder.py:19-250, conditional_single_decoder.py,
conditional_multiple_decoder.py, shared.py + independent.py) on the B200 kernel library.

A base stream (the WACNN of models/wacnn.py) plus a PROGRESSIVE stream: a second analysis transform on
`cat(reshape(y_base), x)` (:226-230,357-358), a second hyperprior and context stack, and per quality level an
importance mask (layers/mask_layer.py) that decides which progressive latents are sent;
`y_hat_q = y_hat_base + y_hat_prog_q` goes through the synthesis transform (one shared decoder for `icd`, a base /
enhancement pair for `imd`).  Same constructor arguments, parameter names and output dictionary as the reference.

The whole multi-quality forward is ONE program: the base stream is computed once (the reference recomputes it
for every quality level, :397-420), the latent-only pre-computations of the progressive context transforms are
shared between levels, and every level's slice loop reuses the kernels / emitters of the base model.
Evaluation mode only (round masks).  `compress` / `decompress` (:510-773) are provided like the base model's: every
tensor the rANS coder consumes is produced on the device (`symbols_and_indexes`, decoder plans), the coder itself
(`compressai.ans`) stays the reference's dependency.
"""
import torch
import torch.nn as nn

from .. import _lib
from ..entropy_models import EntropyBottleneck, GaussianConditional
from ..layers import GDN, GELU, Ctx, Sequential, Win_noShift_Attention, conv, conv3x3, deconv, subpel_conv3x3
from ..layers.base import emit_modules
from ..layers.mask_layer import ONES, ZEROS, Mask
from ..program import TV
from .wacnn import WACNN, SliceDecoder, _cc_stack, _Plan, get_scale_table


class scalable_icd(WACNN):
    """reference scalable/single_decoder.py:25 (registry key "icd")."""

    train_forward_impl = None  # no differentiable training forward for this model: train() gives forward values only
    prog_scale_eps = 0.0       # added to scale_prog * mask before the progressive likelihood (`cimd` only)
    returns_y = True           # forward()'s dict carries "y" (every variant but `cimd`)

    def __init__(self, N=192, M=320, mask_policy="learnable-mask-gamma", lambda_list=(0.05,), lrp_prog=True,
                 independent_lrp=False, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        assert lambda_list is not None
        self.halve = 8
        self.level = 5
        self.factor = self.halve ** 2
        assert N % self.factor == 0
        self.T = N // self.factor + 3
        self.mask_policy = mask_policy
        self.scalable_levels = len(lambda_list)
        self.lmbda_list = list(lambda_list)
        self.lmbda_index_list = dict(zip(self.lmbda_list, range(len(self.lmbda_list))))
        # ---- same construction order as the reference (:59-160): the constructor draws the same init stream
        self.masking = Mask(mask_policy, self.scalable_levels, M)
        self.independent_lrp, self.lrp_prog = independent_lrp, lrp_prog
        if independent_lrp:
            assert lrp_prog is True
            self.lrp_transforms_prog = nn.ModuleList(_cc_stack(320 + 32 * min(i + 1, 6)) for i in range(10))
        self.entropy_bottleneck_prog = EntropyBottleneck(N)
        self.gaussian_conditional_prog = GaussianConditional(None)
        self.g_a_progressive = Sequential(
            conv(self.T, N, kernel_size=5, stride=2), GDN(N),
            conv(N, N, kernel_size=5, stride=2), GDN(N),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            conv(N, N, kernel_size=5, stride=2), GDN(N),
            conv(N, M, kernel_size=5, stride=2))
        self.h_a_prog = Sequential(conv3x3(320, 320), GELU(), conv3x3(320, 288), GELU(), conv3x3(288, 256, stride=2), GELU(),
                                   conv3x3(256, 224), GELU(), conv3x3(224, 192, stride=2))

        def h_s():
            return Sequential(conv3x3(192, 192), GELU(), subpel_conv3x3(192, 224, 2), GELU(), conv3x3(224, 256), GELU(),
                              subpel_conv3x3(256, 288, 2), GELU(), conv3x3(288, 320))

        self.h_mean_s_prog = h_s()
        self.h_scale_s_prog = h_s()
        self.cc_mean_transforms_prog = nn.ModuleList(_cc_stack(320 + 32 * min(i, 5)) for i in range(10))
        self.cc_scale_transforms_prog = nn.ModuleList(_cc_stack(320 + 32 * min(i, 5)) for i in range(10))
        if mask_policy == "learnable-mask":  # (stale policy name kept by the reference, :162-164)
            self.gamma = nn.Parameter(torch.ones((self.scalable_levels - 2, M)))
            self.mask_conv = Sequential(conv(2 * M, M, kernel_size=1, stride=1))

    # ------------------------------------------------------------------ API
    def define_quality(self, quality):
        """reference :232-239"""
        if quality is None:
            return list(self.lmbda_list)
        return list(quality) if isinstance(quality, (list, tuple)) else [quality]

    def _quality_index(self, p):
        return self.lmbda_index_list[p] if p in self.lmbda_index_list else p

    def update(self, scale_table=None, force=False):
        """reference :276-288"""
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated = self.gaussian_conditional_prog.update_scale_table(scale_table, force=force)
        self.entropy_bottleneck_prog.update()
        self.entropy_bottleneck.update()
        return updated

    def _synthesis(self, q):
        return self.g_s

    def _mask_kind(self, q):
        """ZEROS / ONES for the constant masks, None for a computed one."""
        return self.masking.kind(q)

    def _computed_mask(self, ctx, lat_s, lat_sp, q, cache, coding=False):
        """Importance mask of a quality level whose mask is neither all zeros nor all ones (:397-401).  `coding`: the
        call comes from the compress / decompress side (:558,692) rather than from forward."""
        return self.masking.emit(ctx, lat_s, lat_sp, q, cache)

    def _synthesis_input(self, ctx, q, y_hat_q, y_hat_p):
        """What g_s sees at quality index q: the merged latent (as a bf16 copy in bf16 mode)."""
        if ctx.precision == "bf16":
            return ctx.prog.copy(y_hat_q, ctx.buf(y_hat_q.B, y_hat_q.H, y_hat_q.W, self.M))
        return y_hat_q

    def _merge(self, ctx, y_hat, y_hat_p):
        """y_hat_complete = y_hat_base + y_hat_prog (:472)."""
        B, h, w, M = y_hat.B, y_hat.H, y_hat.W, self.M
        return ctx.prog.copy(y_hat, ctx.buf(B, h, w, M, torch.float32), op_code=4, src2=y_hat_p)

    @torch.no_grad()
    def forward(self, x, quality=None, training=True):
        """reference :343-504.  Returns {"x_hat": [Q,B,3,H,W], "likelihoods": {"y": [1,10*B,32,h,w] (slice-major, as the reference's cat over dim 0), "z", "z_prog",
        "y_prog": [Qp,B,M,h,w] (ones if no level uses the progressive stream)}, "y": [Q,B,M,h,w], "z_hat_prog", "z_hat"}."""
        p = self._execute_scalable(x, self.define_quality(quality), False)
        o = self._out
        r = {"x_hat": o(p.x_hat), "likelihoods": {"y": o(p.lik_y), "z": o(p.lik_z), "z_prog": o(p.lik_z_prog), "y_prog": o(p.lik_y_prog)},
             "y": o(p.y_hat_q), "z_hat_prog": o(p.z_hat_prog), "z_hat": o(p.z_hat_out)}
        if not self.returns_y:
            del r["y"]
        return r

    @torch.no_grad()
    def symbols_and_indexes(self, x, quality=0):
        """Everything `compress(x, quality)` (:510-647) computes before the entropy-coder calls, on the device:
        base-stream symbols / CDF indexes / z symbols and, for quality != 0, the progressive stream's masked
        symbols, CDF indexes (of scale_prog * mask) and z_prog symbols."""
        p = self._execute_scalable(x, [quality], True)
        o = self._out
        r = {"y_symbols": o(p.symbols), "y_indexes": o(p.indexes), "z_symbols": o(p.z_symbols), "x_hat": o(p.x_hat[0]),
             "shape": (p.z.H, p.z.W)}
        if p.prog_symbols is not None:
            r.update(y_prog_symbols=o(p.prog_symbols), y_prog_indexes=o(p.prog_indexes), z_prog_symbols=o(p.z_prog_symbols))
        return r

    @torch.no_grad()
    def compress(self, x, quality=0.0):
        """reference :510-647 (every variant inherits or restates it).  ONE program (`symbols_and_indexes`) produces
        every symbol and CDF index of both streams on the device; the rANS coder (`compressai.ans`, imported lazily) is
        called as the reference calls it: one string per image for z / z_prog, one buffered stream for the ten base
        slices of the whole batch, and per progressive slice one string per image (`gaussian_conditional_prog.compress`,
        :613-616).  Returns {"strings": [y_strings, z_strings(, z_prog_strings, progressive_strings)], "shape": [...]}."""
        from ..entropy_models.entropy_models import _EntropyCoder
        cdf, cdf_lengths, offsets = self.gaussian_conditional._coder_tables()
        if self._quality_index(quality) != 0:
            self.gaussian_conditional_prog._coder_tables()  # (fail before the forward pass if update() was not called)
        r = self.symbols_and_indexes(x, quality=quality)
        z_strings = self.entropy_bottleneck.compress(None, symbols=r["z_symbols"])
        c = self.slice_channels
        sym, idx = r["y_symbols"].cpu(), r["y_indexes"].cpu()
        symbols_list, indexes_list = [], []
        for i in range(self.num_slices):
            symbols_list.extend(sym[:, c * i:c * i + c].reshape(-1).tolist())
            indexes_list.extend(idx[:, c * i:c * i + c].reshape(-1).tolist())
        encoder = _EntropyCoder.module().BufferedRansEncoder()
        encoder.encode_with_indexes(symbols_list, indexes_list, cdf, cdf_lengths, offsets)
        y_strings = [encoder.flush()]
        shape = torch.Size(r["shape"])
        if "y_prog_symbols" not in r:
            return {"strings": [y_strings, z_strings], "shape": [shape]}
        z_prog_strings = self.entropy_bottleneck_prog.compress(None, symbols=r["z_prog_symbols"])
        ps, pi = r["y_prog_symbols"].cpu(), r["y_prog_indexes"].cpu()
        progressive = [self.gaussian_conditional_prog._encode_symbols(ps[:, c * i:c * i + c], pi[:, c * i:c * i + c])
                       for i in range(self.num_slices)]
        return {"strings": [y_strings, z_strings, z_prog_strings, progressive], "shape": [shape, shape]}

    @torch.no_grad()
    def decompress(self, strings, shape, quality=None):
        """reference :657-773.  The GPU work of every slice of both streams runs through decoder plans built from the
        same emitters as the forward program (bit-identical CDF indexes); the coder calls are the reference's."""
        from ..entropy_models.entropy_models import _EntropyCoder
        if quality is None:
            raise TypeError("decompress() needs the quality level the strings were coded at")
        q = self._quality_index(quality)
        cdf, cdf_lengths, offsets = self.gaussian_conditional._coder_tables()
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape[0])
        if not z_hat.is_cuda:
            raise RuntimeError("resdsic_b200 runs on CUDA devices only (no CPU fallback)")
        z_hat_prog = self.entropy_bottleneck_prog.decompress(strings[2], shape[-1]) if q != 0 else None
        B, _, hz, wz = z_hat.shape
        key = ("sdec", B, hz, wz, str(z_hat.device), self.precision, q, self._weights_key())
        plan = self._lru_get(self._dec_plans, key, lambda: self._build_scalable_decoder(B, hz, wz, z_hat.device, q), 2)
        run = (lambda prog: prog.run_graph()) if self.use_cuda_graph else (lambda prog: prog.run())
        base = SliceDecoder(self, plan.base, z_hat)
        prog = None
        if q != 0:
            prog = SliceDecoder(self, plan.prog, z_hat_prog)
            if plan.mask_prog is not None:
                run(plan.mask_prog)
        decoder = _EntropyCoder.module().RansDecoder()
        decoder.set_stream(strings[0][0])
        for i in range(self.num_slices):
            idx = base.indexes(i)
            rv = decoder.decode_stream(idx.reshape(-1).tolist(), cdf, cdf_lengths, offsets)
            base.push_symbols(i, torch.tensor(rv, dtype=torch.int32).reshape(idx.shape))
            if prog is not None:
                idx_p = prog.indexes(i)
                prog.push_symbols(i, self.gaussian_conditional_prog._decode_symbols(strings[-1][i], idx_p))
        run(plan.synth)
        return {"x_hat": self._out(plan.x_hat)}

    def _build_scalable_decoder(self, B, hz, wz, device, q, build_only=False):
        """Decoder plans of the base stream and (q != 0) of the progressive stream + the importance mask, and the
        `synth` program: merge (`_merge`), g_s, clamp."""
        from ..program import Program
        ctx = Ctx(device, self.precision, build_only=build_only)
        f32 = torch.float32
        bf16 = ctx.precision == "bf16"
        h, w, M = hz * 4, wz * 4, self.M
        p = _Plan()
        fam = {"cc_mean": self.cc_mean_transforms, "cc_scale": self.cc_scale_transforms, "lrp": self.lrp_transforms}
        p.base = self._emit_decoder_hyper(ctx, B, hz, wz, self.h_mean_s, self.h_scale_s, fam)
        self._emit_decoder_slices(ctx, p.base, fam, self.gaussian_conditional)
        p.prog = p.mask_prog = None
        y_hat = p.base.y_hat
        if q != 0:
            fam_p = {"cc_mean": self.cc_mean_transforms_prog, "cc_scale": self.cc_scale_transforms_prog,
                     "lrp": self.lrp_transforms_prog if self.independent_lrp else self.lrp_transforms}
            p.prog = self._emit_decoder_hyper(ctx, B, hz, wz, self.h_mean_s_prog, self.h_scale_s_prog, fam_p, lrp=self.lrp_prog)
            kind = self._mask_kind(q)
            assert kind != ZEROS
            mask = None
            if kind != ONES:
                ctx.prog = p.mask_prog = Program(device)
                mask = self._computed_mask(ctx, p.base.scales.channels(0, M), p.prog.scales.channels(0, M), q, {}, coding=True)
            self._emit_decoder_slices(ctx, p.prog, fam_p, self.gaussian_conditional_prog, mask=mask, lrp=self.lrp_prog)
        ctx.prog = prog = Program(device)
        if q != 0:
            y_hat = self._merge(ctx, p.base.y_hat, p.prog.y_hat)
        x_raw = torch.empty(B, 3, h * 16, w * 16, dtype=f32, device=device)
        p.x_hat = torch.empty(B, 3, h * 16, w * 16, dtype=f32, device=device)
        act = self._synthesis_input(ctx, q, y_hat, p.prog.y_hat if q != 0 else None)
        self._synthesis(q).emit(ctx, act, last_kw=dict(out=TV.nchw_of(x_raw)))
        prog.copy(TV.nchw_of(x_raw), TV.nchw_of(p.x_hat), op_code=3)  # clamp_(0, 1), :771
        p.synth = prog
        p.y_hat_q = y_hat
        return p

    # ------------------------------------------------------------- planning
    def _execute_scalable(self, x, qualities, with_symbols):
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected [B,3,H,W] input, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("resdsic_b200 runs on CUDA devices only (no CPU fallback)")
        if self.training:
            raise RuntimeError("the scalable models are evaluation-only in resdsic_b200 (call .eval())")
        B, _, H, W = x.shape
        if H % 64 or W % 64:
            raise ValueError(f"input {H}x{W} must be a multiple of 64")
        qs = tuple(self._quality_index(q) for q in qualities)
        key = ("scalable", B, H, W, str(x.device), self.precision, with_symbols, qs, self._weights_key())
        plan = self._lru_get(self._plans, key, lambda: self._build_scalable(B, H, W, x.device, qs, with_symbols), self.MAX_PLANS)
        self._last_plan = plan
        plan.x.copy_(x)
        plan.prog.run_graph() if self.use_cuda_graph else plan.prog.run()
        self.last_num_launches = plan.prog.num_launches
        return plan

    def _build_scalable(self, B, H, W, device, qs, with_symbols, build_only=False):
        ctx = Ctx(device, self.precision, build_only=build_only)
        f32, i32 = torch.float32, torch.int32
        bf16 = ctx.precision == "bf16"
        prog = ctx.prog
        p = _Plan()
        h, w = H // 16, W // 16
        M, N = self.M, self.N
        new = lambda *shape, dtype=f32: torch.empty(*shape, dtype=dtype, device=device)
        p.x = new(B, 3, H, W)
        # ---- g_a in two parts (split_ga, :196-202): y_base = g_a[:6](x) (before the third GDN), y = g_a[6:](y_base)
        mods = list(self.g_a)
        y_base = emit_modules(ctx, mods[: self.level + 1], TV.nchw_of(p.x))
        y = ctx.buf(B, h, w, M, f32)
        y_act = ctx.buf(B, h, w, M) if bf16 else y
        emit_modules(ctx, mods[self.level + 1:], y_base, last_kw=dict(out=y, out2=y_act) if bf16 else dict(out=y))
        # ---- concatenate (:226-230): the NCHW memory of y_base [B,N,H/8,W/8] REINTERPRETED as [B,N/64,H,W], cat x
        xp = new(B, self.T, H, W)
        prog.copy(y_base, TV.nchw_channels(xp, 0, N, H=H // 8, W=W // 8, total=self.T * self.factor))
        prog.copy(TV.nchw_of(p.x), TV.nchw_channels(xp, self.T - 3, 3))
        y_prog = ctx.buf(B, h, w, M, f32)
        y_prog_act = ctx.buf(B, h, w, M) if bf16 else y_prog
        self.g_a_progressive.emit(ctx, TV.nchw_of(xp), last_kw=dict(out=y_prog, out2=y_prog_act) if bf16 else dict(out=y_prog))
        # ---- the two hyperpriors
        hz, wz = h // 4, w // 4
        p.lik_z, p.lik_z_prog = new(B, N, hz, wz), new(B, N, hz, wz)
        p.z_symbols = new(B, N, hz, wz, dtype=i32) if with_symbols else None
        p.z_prog_symbols = new(B, N, hz, wz, dtype=i32) if with_symbols else None
        z, z_hat, means, scales = self._emit_hyperprior(ctx, y_act, self.h_a, self.entropy_bottleneck, self.h_mean_s,
                                                        self.h_scale_s, p.lik_z, p.z_symbols)
        z_p, z_hat_p, means_p, scales_p = self._emit_hyperprior(ctx, y_prog_act, self.h_a_prog, self.entropy_bottleneck_prog,
                                                                self.h_mean_s_prog, self.h_scale_s_prog, p.lik_z_prog,
                                                                p.z_prog_symbols)
        p.z_hat_out, p.z_hat_prog = new(B, N, hz, wz), new(B, N, hz, wz)
        prog.copy(z_hat, TV.nchw_of(p.z_hat_out))
        prog.copy(z_hat_p, TV.nchw_of(p.z_hat_prog))
        # ---- base slice loop, ONCE (identical for every quality level in the reference)
        # forward(): likelihoods["y"] is [1, 10*B, 32, h, w], slice-major -- the reference concatenates the ten slice
        # likelihoods along dim 0 (:480); with_symbols keeps the standard [B,M,h,w] layout next to symbols / indexes
        p.lik_y = new(B, M, h, w) if with_symbols else new(1, self.num_slices * B, self.slice_channels, h, w)
        p.symbols = new(B, M, h, w, dtype=i32) if with_symbols else None
        p.indexes = new(B, M, h, w, dtype=i32) if with_symbols else None
        fam = {"cc_mean": self.cc_mean_transforms, "cc_scale": self.cc_scale_transforms, "lrp": self.lrp_transforms}
        pre = self._emit_slice_precompute(ctx, fam, means, scales)
        y_hat = self._emit_slice_loop(ctx, fam, pre, self.gaussian_conditional, y, means, scales, p.lik_y, p.symbols, p.indexes,
                                      lik_slice_major=not with_symbols)
        # ---- progressive stream per quality level
        prog_qs = [q for q in qs if q != 0]
        p.lik_y_prog = new(len(prog_qs), B, M, h, w) if prog_qs else torch.ones(1, B, M, h, w, dtype=f32, device=device)
        p.prog_symbols = new(B, M, h, w, dtype=i32) if (with_symbols and prog_qs) else None
        p.prog_indexes = new(B, M, h, w, dtype=i32) if (with_symbols and prog_qs) else None
        fam_p = {"cc_mean": self.cc_mean_transforms_prog, "cc_scale": self.cc_scale_transforms_prog,
                 "lrp": self.lrp_transforms_prog if self.independent_lrp else self.lrp_transforms}
        pre_p = self._emit_slice_precompute(ctx, fam_p, means_p, scales_p) if prog_qs else None
        lat_s, lat_sp = scales.channels(0, M), scales_p.channels(0, M)
        mask_cache, first = {}, True
        p.x_hat = new(len(qs), B, 3, H, W)
        p.y_hat_q = new(len(qs), B, M, h, w)
        p.masks = {}
        jp = 0
        for j, q in enumerate(qs):
            y_hat_q = y_hat
            if q != 0:
                kind = self._mask_kind(q)
                assert kind != ZEROS
                mask = None if kind == ONES else self._computed_mask(ctx, lat_s, lat_sp, q, mask_cache, coding=with_symbols)
                p.masks[q] = mask
                if first:
                    mq, sq = means_p, scales_p
                else:  # later levels: private support slots over a copy of the progressive latents
                    mq, sq = ctx.buf(B, h, w, means_p.ld), ctx.buf(B, h, w, scales_p.ld)
                    prog.copy(means_p.channels(0, M), mq.channels(0, M))
                    prog.copy(scales_p.channels(0, M), sq.channels(0, M))
                first = False
                y_hat_p = self._emit_slice_loop(ctx, fam_p, pre_p, self.gaussian_conditional_prog, y_prog, mq, sq,
                                                p.lik_y_prog[jp], p.prog_symbols, p.prog_indexes, mask=mask, lrp=self.lrp_prog,
                                                # compress() (single_decoder.py:583, inherited by every variant) indexes scale * mask
                                                scale_eps=0.0 if with_symbols else self.prog_scale_eps)
                jp += 1
                y_hat_q = self._merge(ctx, y_hat, y_hat_p)
            else:
                y_hat_p = None
            prog.copy(y_hat_q, TV.nchw_of(p.y_hat_q[j]))
            act = self._synthesis_input(ctx, q, y_hat_q, y_hat_p)
            self._synthesis(q).emit(ctx, act, last_kw=dict(out=TV.nchw_of(p.x_hat[j])))
        p.prog = prog
        p.y, p.z, p.y_hat, p.means, p.scales, p.y_prog, p.y_base = y, z, y_hat, means, scales, y_prog, y_base
        p.subs, p.sub_batch = [p], B
        return p


class scalable_imd(scalable_icd):
    """reference scalable/multiple_decoder.py:19 (registry key "imd"): `icd` with a base decoder g_s[0]
    (quality 0) and an enhancement decoder g_s[1] (:36-50,225)."""

    def __init__(self, N=192, M=320, mask_policy="learnable-mask-gamma", lambda_list=(0.05,), lrp_prog=True,
                 independent_lrp=False, **kwargs):
        super().__init__(N=N, M=M, mask_policy=mask_policy, lambda_list=lambda_list, lrp_prog=lrp_prog,
                         independent_lrp=independent_lrp, **kwargs)
        self.g_s = nn.ModuleList(
            Sequential(
                Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2),
                deconv(M, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
                deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                deconv(N, 3, kernel_size=5, stride=2)) for _ in range(2))

    def _synthesis(self, q):
        return self.g_s[0 if q == 0 else 1]


class conditional_scalable_icd(scalable_icd):
    """reference scalable/conditional_single_decoder.py:18 (registry key "cicd"): `icd` whose base and progressive
    reconstructions of a slice are MERGED by a policy instead of summed (:103-113) -- "conditional": a per-slice
    joiner network conv3x3(64,64)-GELU-conv3x3(64,64)-GELU-conv3x3(64,32) on cat(y_hat_slice, y_hat_prog_slice)
    (:39-48); "residual": the sum; "concatenation" / "cac": the base reconstruction alone.  The reference calls
    `self.masking(latent_scales, pr=quality)` WITHOUT the progressive scales (:163), so only the constant masks are
    reachable (two-levels, or the end points of the learnable policies: anything else trips Mask.forward's
    `assert scale_prog is not None`, mask_layer.py:74,98), and with an all-ones mask its unmasked likelihood /
    masked reconstruction (:221-225) coincide with `icd`'s.  Like the reference's, the default constructor arguments
    (`mask_policy="learnable-mask"`, one lambda) do not construct (`gamma` of -1 rows, single_decoder.py:162-164)."""

    def __init__(self, N=192, M=320, mask_policy="learnable-mask", lambda_list=(0.05,), lrp_prog=True,
                 independent_lrp=False, joiner_policy="conditional", **kwargs):
        super().__init__(N=N, M=M, mask_policy=mask_policy, lambda_list=lambda_list, lrp_prog=lrp_prog,
                         independent_lrp=independent_lrp, **kwargs)
        self.joiner_policy = joiner_policy
        if joiner_policy == "conditional":
            self.joiner = nn.ModuleList(
                Sequential(conv(64, 64, stride=1, kernel_size=3), GELU(), conv(64, 64, stride=1, kernel_size=3), GELU(),
                           conv(64, 32, stride=1, kernel_size=3)) for _ in range(10))
        elif joiner_policy == "cac":  # (the reference overwrites the policy string with this module, :49-54; kept as a module here)
            self.joiner_cac = conv(M, M, kernel_size=1, stride=1)
        elif joiner_policy not in ("residual", "concatenation"):
            raise NotImplementedError(f"joiner policy {joiner_policy!r} (block_concatenation changes the decoder width: not provided)")

    def _computed_mask(self, ctx, lat_s, lat_sp, q, cache, coding=False):
        if coding:  # compress / decompress are `icd`'s (inherited): they do pass the progressive scales (:558,692)
            return super()._computed_mask(ctx, lat_s, lat_sp, q, cache, coding)
        raise AssertionError("scale_prog is None: the reference's cicd forward reaches only all-zero / all-one masks "
                             "(conditional_single_decoder.py:163, mask_layer.py:74,98)")

    def _merge(self, ctx, y_hat, y_hat_p):
        if self.joiner_policy == "residual":
            return super()._merge(ctx, y_hat, y_hat_p)
        if self.joiner_policy in ("concatenation", "cac"):
            return y_hat
        from ..layers.conv import emit_grouped
        B, h, w, M, sc_ = y_hat.B, y_hat.H, y_hat.W, self.M, self.slice_channels
        f32 = torch.float32
        bf16 = ctx.precision == "bf16"
        stacks = [[m for m in seq if hasattr(m, "weight")] for seq in self.joiner]
        a = ctx.prog.copy(y_hat, ctx.buf(B, h, w, M)) if bf16 else y_hat      # A operands of the joiner's first conv
        b = ctx.prog.copy(y_hat_p, ctx.buf(B, h, w, M)) if bf16 else y_hat_p
        out = ctx.buf(B, h, w, M, f32)
        if bf16:
            # all ten joiners as grouped launches; conv(cat(main_i, prog_i); W) = conv(main_i; W[:, :32]) + conv(prog_i; W[:, 32:])
            first = [cs[0] for cs in stacks]
            part = emit_grouped(ctx, self, "join0a", first, a.channels(0, sc_), sc_, cols=(0, sc_), bias=False, out_dtype=f32)
            t = emit_grouped(ctx, self, "join0b", first, b.channels(0, sc_), sc_, cols=(sc_, 2 * sc_), res=part, gelu=True)
            t = emit_grouped(ctx, self, "join1", [cs[1] for cs in stacks], t.channels(0, 64), 64, gelu=True)
            emit_grouped(ctx, self, "join2", [cs[2] for cs in stacks], t.channels(0, 64), 64, out=out)
            return out
        for i, cs in enumerate(stacks):
            part = emit_grouped(ctx, self, ("join0a", i), [cs[0]], a.channels(sc_ * i, sc_), 0, cols=(0, sc_), bias=False, out_dtype=f32)
            t = emit_grouped(ctx, self, ("join0b", i), [cs[0]], b.channels(sc_ * i, sc_), 0, cols=(sc_, 2 * sc_), res=part, gelu=True)
            t = cs[1].emit(ctx, t, gelu=True)
            cs[2].emit(ctx, t, out=out.channels(sc_ * i, sc_))
        return out


def _synthesis_pair(M, N, dims=None):
    """Base / enhancement decoder pair (multiple_decoder.py:36-50); `dims`: input width of each decoder."""
    dims = dims or [M, M]
    return nn.ModuleList(
        Sequential(
            Win_noShift_Attention(dim=dims[i], num_heads=8, window_size=4, shift_size=2),
            deconv(dims[i], N, kernel_size=5, stride=2), GDN(N, inverse=True),
            deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
            deconv(N, 3, kernel_size=5, stride=2)) for i in range(2))


class conditional_scalable_imd(conditional_scalable_icd):
    """reference scalable/conditional_multiple_decoder.py:20 (registry key "cimd"): `cicd` with a base / enhancement
    decoder pair (:43-55,236).  Unlike `cicd`, its forward hands the progressive scales to `Mask.forward` (:158), so the
    computed masks of the learnable policies ARE reachable; the progressive likelihood is taken at
    `scale_prog * mask + 1e-7` (:210) and the forward dictionary has no "y" entry (:262-267).  compress() is `icd`'s
    (inherited in the reference), so `symbols_and_indexes` indexes `scale_prog * mask` without the 1e-7.
    `joiner_policy="concatenation"`: the enhancement decoder g_s[1] is 2M wide and is fed with
    cat(base reconstruction, progressive reconstruction) (:41,230; `merge` returns the base alone for this policy)."""

    prog_scale_eps = 1e-7
    returns_y = False

    def __init__(self, N=192, M=320, mask_policy="learnable-mask", lambda_list=(0.05,), lrp_prog=True,
                 independent_lrp=False, joiner_policy="conditional", **kwargs):
        super().__init__(N=N, M=M, mask_policy=mask_policy, lambda_list=lambda_list, lrp_prog=lrp_prog,
                         independent_lrp=independent_lrp, joiner_policy=joiner_policy, **kwargs)
        self.dimensions_M = [M, 2 * M if joiner_policy == "concatenation" else M]
        self.g_s = _synthesis_pair(M, N, self.dimensions_M)

    def _synthesis_input(self, ctx, q, y_hat_q, y_hat_p):
        if self.joiner_policy != "concatenation" or q == 0:
            return super()._synthesis_input(ctx, q, y_hat_q, y_hat_p)
        B, h, w, M = y_hat_q.B, y_hat_q.H, y_hat_q.W, self.M  # cat(y_hat_complete (= base), y_hat_prog), :230
        cat = ctx.buf(B, h, w, 2 * M) if ctx.precision == "bf16" else ctx.buf(B, h, w, 2 * M, torch.float32)
        ctx.prog.copy(y_hat_q, cat.channels(0, M))
        ctx.prog.copy(y_hat_p, cat.channels(M, M))
        return cat

    def _synthesis(self, q):
        return self.g_s[0 if q == 0 else 1]

    def _computed_mask(self, ctx, lat_s, lat_sp, q, cache, coding=False):
        return scalable_icd._computed_mask(self, ctx, lat_s, lat_sp, q, cache, coding)


class ResWACNNIndependentEntropy(scalable_icd):
    """reference scalable/shared.py:23 + scalable/independent.py:24 (registry key "ind"): base and progressive streams
    with independent hyperpriors / context stacks and NO importance mask on the data path -- `extract_mask`
    (shared.py:191-229) is evaluated by the reference's forward and compress but its result is never applied
    (independent.py:318-392,545-580: the `block_mask` lines are commented out), so every quality index != 0 sends the
    whole progressive stream.  Parameters follow the reference's names and construction order (`gamma` / `mask_conv`
    of the "learnable-mask" policy live on the model itself, shared.py:69-71, and stay unused); `multiple_decoder`
    selects a base / enhancement decoder pair (independent.py:129-143,438-441)."""

    MASK_POLICIES = ("point-based-std", "learnable-mask", "all-one", "all-zero", "two-levels")

    def __init__(self, N=192, M=320, mask_policy="two-levels", lambda_list=(0.0035, 0.065), lrp_prog=True,
                 independent_lrp=False, multiple_decoder=False, **kwargs):
        WACNN.__init__(self, N=N, M=M, **kwargs)
        assert lambda_list is not None
        self.halve, self.level = 8, 5
        self.factor = self.halve ** 2
        assert N % self.factor == 0
        self.T = N // self.factor + 3
        self.mask_policy = mask_policy
        self.scalable_levels = len(lambda_list)
        self.lmbda_list = list(lambda_list)
        self.lmbda_index_list = dict(zip(self.lmbda_list, range(len(self.lmbda_list))))
        # ---- shared.py:54-71 (the constructor draws the reference's init stream in the reference's order)
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.entropy_bottleneck_prog = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)
        self.gaussian_conditional_prog = GaussianConditional(None)
        self.g_a_progressive = Sequential(
            conv(self.T, N, kernel_size=5, stride=2), GDN(N),
            conv(N, N, kernel_size=5, stride=2), GDN(N),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            conv(N, N, kernel_size=5, stride=2), GDN(N),
            conv(N, M, kernel_size=5, stride=2))
        if mask_policy == "learnable-mask":
            self.gamma = nn.Parameter(torch.ones((self.scalable_levels - 1, M)))
            self.mask_conv = Sequential(conv(M, M, kernel_size=1, stride=1))
        # ---- independent.py:42-127
        self.multiple_decoder = multiple_decoder
        self.h_a_prog = Sequential(conv3x3(320, 320), GELU(), conv3x3(320, 288), GELU(), conv3x3(288, 256, stride=2), GELU(),
                                   conv3x3(256, 224), GELU(), conv3x3(224, 192, stride=2))

        def h_s():
            return Sequential(conv3x3(192, 192), GELU(), subpel_conv3x3(192, 224, 2), GELU(), conv3x3(224, 256), GELU(),
                              subpel_conv3x3(256, 288, 2), GELU(), conv3x3(288, 320))

        self.h_mean_s_prog = h_s()
        self.h_scale_s_prog = h_s()
        self.cc_mean_transforms_prog = nn.ModuleList(_cc_stack(320 + 32 * min(i, 5)) for i in range(10))
        self.cc_scale_transforms_prog = nn.ModuleList(_cc_stack(320 + 32 * min(i, 5)) for i in range(10))
        self.independent_lrp = independent_lrp
        if independent_lrp:
            self.lrp_transforms_prog = nn.ModuleList(_cc_stack(320 + 32 * min(i + 1, 6)) for i in range(10))
        self.lrp_prog = lrp_prog
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.entropy_bottleneck_prog = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)
        self.gaussian_conditional_prog = GaussianConditional(None)
        if multiple_decoder:
            self.g_s = _synthesis_pair(M, N)

    def _mask_kind(self, q):
        if self.mask_policy not in self.MASK_POLICIES:
            raise NotImplementedError(self.mask_policy)  # shared.py:229
        return ZEROS if q == 0 else ONES

    def _synthesis(self, q):
        return self.g_s[0 if q == 0 else 1] if self.multiple_decoder else self.g_s
