"""EntropyBottleneck / GaussianConditional on the fused CUDA entropy kernels.

Mirror of the reference's vendored CompressAI classes
(entropy_models/entropy_models.py:295-668): same constructor arguments,
parameter/buffer names and forward semantics in eval mode (round) and in
training mode (additive U(-1/2,1/2) noise, forward VALUES only -- there is no
autograd through the CUDA kernels).  In scope: forward (quantise + likelihood),
quantize, dequantize, build_indexes, EntropyBottleneck.loss and the CDF-table build
`update()`, and the compress()/decompress() glue around the rANS coder.  Out of
scope per BASELINE.json (stays in the reference's C++): the rANS coder itself, `compressai.ans`,
imported lazily like the reference does.
"""
import ctypes

from .. import _lib
import numpy as np
import torch
import torch.nn as nn

from .. import packing
from ..layers.base import B200Module, Ctx
from ..ops import LowerBound
from ..program import TV


class _EntropyCoder:
    """Proxy to the rANS coder of pip CompressAI (`compressai.ans`, C++), exactly as the reference's
    `_EntropyCoder` (entropy_models.py:16-50) resolves it.  The coder itself is out of scope of this package
    (BASELINE.json: "rANS bitstream coding stays in the reference's C++"): it is imported lazily, so everything
    else works without it and the bitstream entry points fail with a clear message when it is absent."""

    def __init__(self, method="ans"):
        if method != "ans":
            raise ValueError(f'Unknown entropy coder "{method}" (available: ans)')
        self.name = method
        self._encoder = self._decoder = None

    @staticmethod
    def module():
        try:
            from compressai import ans
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("the rANS coder (`compressai.ans` of pip CompressAI, C++) is not importable here; "
                               "compress()/decompress() need it -- every tensor they exchange with it is available "
                               "through symbols_and_indexes() / update() / slice_decoder()") from e
        return ans

    def encode_with_indexes(self, *args, **kwargs):
        if self._encoder is None:
            self._encoder = self.module().RansEncoder()
        return self._encoder.encode_with_indexes(*args, **kwargs)

    def decode_with_indexes(self, *args, **kwargs):
        if self._decoder is None:
            self._decoder = self.module().RansDecoder()
        return self._decoder.decode_with_indexes(*args, **kwargs)


class EntropyModel(B200Module):
    """reference entropy_models.py:70-292 (buffers only; coding is out of scope)."""

    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder = _EntropyCoder(entropy_coder or "ans")
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        self.likelihood_bound = float(likelihood_bound)
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    # -- CDF tables (update()): what the reference's rANS coder consumes; built on the device --------------
    @staticmethod
    def _stream(dev):
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    @staticmethod
    def _require_cuda(t, what):
        if t.device.type != "cuda":
            raise RuntimeError(f"resdsic_b200 {what} runs on CUDA devices only (no CPU fallback)")

    def _pmf_to_cdf(self, prob, cdf_length, max_length):
        """reference entropy_models.py:174-182 + `pmf_to_quantized_cdf`: prob [rows, max_length+1] fp32 holds
        pmf[:pmf_length] ++ tail_mass per row; returns int32 [rows, max_length+2], zero padded."""
        dev = prob.device
        rows = prob.shape[0]
        cdf = torch.empty(rows, max_length + 2, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rdsic_pmf_to_quantized_cdf(prob.data_ptr(), prob.shape[1], cdf_length.data_ptr(), rows,
                                                             self.entropy_coder_precision, cdf.data_ptr(), cdf.shape[1],
                                                             status.data_ptr(), self._stream(dev)), "pmf_to_quantized_cdf")
        bad = int(status.item())
        if bad:  # reference: std::domain_error from the C++ op -> ValueError-like failure in update()
            raise ValueError(f"Invalid `pmf` in table row {bad - 1}: negative, non-finite or all-zero probabilities")
        return cdf

    def _check_cdf_size(self):
        """reference entropy_models.py:184-203."""
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def _coder_tables(self):
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        return (self._quantized_cdf.tolist(), self._cdf_length.reshape(-1).int().tolist(),
                self._offset.reshape(-1).int().tolist())

    def _encode_symbols(self, symbols, indexes):
        """reference entropy_models.py:227-238: one string per batch element (`symbols`/`indexes` int32 [B,...])."""
        if symbols.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if symbols.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        cdf, lengths, offsets = self._coder_tables()
        symbols, indexes = symbols.cpu(), indexes.cpu()
        return [self.entropy_coder.encode_with_indexes(symbols[i].reshape(-1).int().tolist(),
                                                       indexes[i].reshape(-1).int().tolist(), cdf, lengths, offsets)
                for i in range(symbols.size(0))]

    def _decode_symbols(self, strings, indexes):
        """reference entropy_models.py:241-287 up to the dequantisation: int32 symbols shaped like `indexes`."""
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        cdf, lengths, offsets = self._coder_tables()
        idx = indexes.cpu()
        out = torch.empty(idx.size(), dtype=torch.int32)
        for i, st in enumerate(strings):
            values = self.entropy_coder.decode_with_indexes(st, idx[i].reshape(-1).int().tolist(), cdf, lengths, offsets)
            out[i] = torch.tensor(values, dtype=torch.int32).reshape(out[i].size())
        return out

    @staticmethod
    def dequantize(inputs, means=None):
        """reference entropy_models.py:160-167 (pure dtype/add glue, used after rANS decode)."""
        if means is not None:
            return inputs.type_as(means) + means
        return inputs.float()


class EntropyBottleneck(EntropyModel):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, filters[i + 1], filters[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, filters[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, filters[i + 1], 1)))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles = nn.Parameter(init.repeat(channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def packed(self):
        names = [f"_matrix{i}" for i in range(5)] + [f"_bias{i}" for i in range(5)] + [f"_factor{i}" for i in range(4)]
        tensors = [getattr(self, n) for n in names] + [self.quantiles]

        def build():
            # softplus/tanh of the parameters on the CPU: the same ATen ops the reference applies
            cpu = {n: getattr(self, n).detach().cpu() for n in names}
            return packing.pack_entropy_bottleneck(cpu, self.quantiles.detach().cpu()).to(self.quantiles.device)

        return self._packed("eb", tensors, build)

    def emit(self, ctx: Ctx, z, z_hat=None, lik=None, symbols=None, noise=None, noisy_out=None, **kw):
        """z: fp32 [B,h,w,C] view.  Returns (z_hat view, lik NCHW tensor).  `noise` (fp32 view like z)
        selects the training-mode likelihood at z + noise; z_hat stays round(z-med)+med (cnn.py:152-154)."""
        if z_hat is None:
            z_hat = ctx.buf(z.B, z.H, z.W, z.C)
        if lik is None:
            lik = torch.empty(z.B, z.C, z.H, z.W, dtype=torch.float32, device=ctx.device)
        ctx.prog.eb(z, z_hat, lik, self.packed(), symbols=symbols, lik_bound=self.likelihood_bound, noise=noise,
                    noisy_out=noisy_out)
        return z_hat, lik

    @torch.no_grad()
    def forward(self, x, training=None, noise=None):
        """(outputs, likelihood) -- reference entropy_models.py:447-490.  Eval: outputs = round(x-med)+med.
        Training: outputs = x + U(-1/2,1/2) (`noise`, NCHW, may be injected for reproducibility; otherwise it
        is drawn on the device with torch's generator) and the likelihood is evaluated there."""
        training = self.training if training is None else training
        ctx = Ctx(x.device, "fp32")
        z = ctx.from_nchw(x, torch.float32)
        nz = noisy = None
        if training:
            if noise is None:
                noise = torch.empty_like(x, dtype=torch.float32).uniform_(-0.5, 0.5)
            nz = ctx.from_nchw(noise, torch.float32)
            noisy = ctx.buf(z.B, z.H, z.W, z.C, torch.float32)
        z_hat, lik = self.emit(ctx, z, z_hat=ctx.buf(z.B, z.H, z.W, z.C, torch.float32), noise=nz, noisy_out=noisy)
        out = ctx.to_nchw(noisy if training else z_hat)
        ctx.prog.run()
        return out, lik

    @torch.no_grad()
    def update(self, force=False):
        """reference entropy_models.py:356-394 (this fork recomputes unconditionally; `force` is accepted and
        ignored like there).  Fills `_offset`, `_quantized_cdf`, `_cdf_length` on the parameters' device."""
        q = self.quantiles.detach().float().contiguous()
        dev = q.device
        self._require_cuda(q, "EntropyBottleneck.update")
        Cn = self.channels
        L = _lib.lib()
        offset = torch.empty(Cn, dtype=torch.int32, device=dev)
        cdf_length = torch.empty(Cn, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.rdsic_eb_cdf_sizes(q.data_ptr(), Cn, offset.data_ptr(), cdf_length.data_ptr(), self._stream(dev)),
                       "EntropyBottleneck.update sizes")
            max_length = int(cdf_length.max().item()) - 2  # the one host sync: the table width
            prob = torch.zeros(Cn, max_length + 1, dtype=torch.float32, device=dev)
            _lib.check(L.rdsic_eb_pmf(self.packed().data_ptr(), q.data_ptr(), offset.data_ptr(), cdf_length.data_ptr(), Cn,
                                      max_length, prob.data_ptr(), prob.shape[1], self._stream(dev)),
                       "EntropyBottleneck.update pmf")
        self._quantized_cdf = self._pmf_to_cdf(prob, cdf_length, max_length)
        self._offset, self._cdf_length = offset, cdf_length
        self.__dict__["_last_prob"] = prob  # kept for the parity tests (float stage)
        return True

    @torch.no_grad()
    def loss(self):
        """reference entropy_models.py:396-399: sum |logits_cumulative(quantiles) - target| (forward value)."""
        dev = self.quantiles.device
        if dev.type != "cuda":
            raise RuntimeError("resdsic_b200 runs on CUDA devices only (no CPU fallback)")
        q = self.quantiles.detach().float().contiguous()
        tgt = self.target.detach().float().contiguous()
        out = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().rdsic_eb_aux_loss(self.packed().data_ptr(), q.data_ptr(), tgt.data_ptr(), self.channels, None,
                                              out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "EntropyBottleneck.loss")
        return out[0]

    @staticmethod
    def _build_indexes(size):
        """reference entropy_models.py:492-503: index = channel id."""
        N, Cn = size[0], size[1]
        idx = torch.arange(Cn, dtype=torch.int32).view(1, -1, *([1] * (len(size) - 2)))
        return idx.repeat(N, 1, *size[2:])

    @torch.no_grad()
    def symbols(self, x):
        """round(x - median) as int32 NCHW (quantize "symbols" with the medians, entropy_models.py:139-152,512-519)."""
        ctx = Ctx(x.device, "fp32")
        z = ctx.from_nchw(x, torch.float32)
        sym = torch.empty(z.B, z.C, z.H, z.W, dtype=torch.int32, device=x.device)
        self.emit(ctx, z, z_hat=ctx.buf(z.B, z.H, z.W, z.C, torch.float32), symbols=sym)
        ctx.prog.run()
        return sym

    @torch.no_grad()
    def compress(self, x, symbols=None):
        """reference entropy_models.py:512-519: one rANS string per batch element.  `symbols` (int32 NCHW) may be
        passed when a forward pass has already produced them."""
        if symbols is None:
            symbols = self.symbols(x)
        return self._encode_symbols(symbols, self._build_indexes(symbols.size()))

    @torch.no_grad()
    def decompress(self, strings, size):
        """reference entropy_models.py:521-526: decoded symbols + medians, float [B,C,*size] on the module's device."""
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        sym = self._decode_symbols(strings, self._build_indexes(output_size))
        med = self._get_medians().detach().reshape(1, -1, *([1] * len(size)))
        return self.dequantize(sym.to(med.device), med)


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.scale_bound_value = float(scale_bound)
        self.register_buffer("scale_table", torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _prepare_scale_table(scale_table):
        """reference entropy_models.py:575-577."""
        return torch.Tensor(tuple(float(x) for x in scale_table))

    @staticmethod
    def _standardized_quantile(quantile):
        """reference entropy_models.py:586-588 (host scalar)."""
        import scipy.stats
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        """reference entropy_models.py:590-598."""
        device = self.scale_bound.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    @torch.no_grad()
    def update(self):
        """reference entropy_models.py:599-625: one CDF row per scale-table entry."""
        table = self.scale_table.detach().float().contiguous()
        dev = table.device
        self._require_cuda(table, "GaussianConditional.update")
        if table.numel() == 0:
            raise ValueError("empty scale_table: call update_scale_table(scale_table) first")
        n = table.numel()
        multiplier = float(-self._standardized_quantile(self.tail_mass / 2))
        L = _lib.lib()
        offset = torch.empty(n, dtype=torch.int32, device=dev)
        cdf_length = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.rdsic_gc_cdf_sizes(table.data_ptr(), n, multiplier, offset.data_ptr(), cdf_length.data_ptr(),
                                            self._stream(dev)), "GaussianConditional.update sizes")
            max_length = int(cdf_length.max().item()) - 2  # the one host sync: the table width
            prob = torch.zeros(n, max_length + 1, dtype=torch.float32, device=dev)
            _lib.check(L.rdsic_gc_pmf(table.data_ptr(), offset.data_ptr(), n, prob.data_ptr(), prob.shape[1],
                                      self._stream(dev)), "GaussianConditional.update pmf")
        self._quantized_cdf = self._pmf_to_cdf(prob, cdf_length, max_length)
        self._offset, self._cdf_length = offset, cdf_length
        self.__dict__["_last_prob"] = prob
        self.__dict__.pop("_pack_cache", None)  # the device copy of the scale table follows the new buffer

    def table(self, device):
        """Device copy of the scale table; falls back to the model default
        (reference cnn.py:14-20) when update() has not filled the buffer."""
        t = self.scale_table
        if t.numel() == 0:
            t = self.__dict__.setdefault("_default_table", packing.scale_table())
        return self._packed(("table", str(device)), (t,), lambda: t.detach().float().to(device).contiguous())

    def emit(self, ctx: Ctx, y, scale, mu, lik, lik_coff, Ctot, y_hat_dsts=(), symbols=None, indexes=None, noise=None,
             noisy_out=None, sym_in=None, mask=None, scale_eps=0.0):
        """`noise` (fp32 view like y) selects the training-mode likelihood at y + noise (entropy_models.py:646-661);
        the y_hat destinations keep the ste_round value (cnn.py:177).  `mask` (fp32 view like y): the ResDSIC
        progressive stream (scalable/single_decoder.py:447-453); `scale_eps`: added to scale * mask before the
        likelihood (`cimd`, scalable/conditional_multiple_decoder.py:210)."""
        ctx.prog.gc(y, mu, scale, list(y_hat_dsts), lik, lik_coff, Ctot, self.table(ctx.device), symbols=symbols,
                    indexes=indexes, scale_bound=self.scale_bound_value,
                    lik_bound=self.likelihood_bound, noise=noise, noisy_out=noisy_out, sym_in=sym_in, mask=mask,
                    scale_eps=scale_eps)

    def _run(self, inputs, scales, means, want, noise=None):
        B, Cn, H, W = inputs.shape
        ctx = Ctx(inputs.device, "fp32")
        y = ctx.from_nchw(inputs, torch.float32)
        sc = ctx.from_nchw(scales, torch.float32)
        mu = ctx.from_nchw(means if means is not None else torch.zeros_like(inputs), torch.float32)
        y_hat = ctx.buf(B, H, W, Cn, torch.float32)
        lik = torch.empty(B, Cn, H, W, dtype=torch.float32, device=inputs.device)
        sym = torch.empty(B, Cn, H, W, dtype=torch.int32, device=inputs.device) if "sym" in want else None
        idx = torch.empty(B, Cn, H, W, dtype=torch.int32, device=inputs.device) if "idx" in want else None
        nz = noisy = None
        if noise is not None:
            nz = ctx.from_nchw(noise, torch.float32)
            noisy = ctx.buf(B, H, W, Cn, torch.float32)
        self.emit(ctx, y, sc, mu, lik, 0, Cn, [y_hat], sym, idx, noise=nz, noisy_out=noisy)
        out = ctx.to_nchw(noisy if noise is not None else y_hat)
        ctx.prog.run()
        return out, lik, sym, idx

    @staticmethod
    def _draw_noise(inputs, mask):
        """quantize "noise" (entropy_models.py:131-137): U(-1/2,1/2), optionally times `mask`."""
        noise = torch.empty_like(inputs, dtype=torch.float32).uniform_(-0.5, 0.5)
        return noise if mask is None else noise * mask

    @torch.no_grad()
    def forward(self, inputs, scales, means=None, training=None, mask=None, noise=None):
        """(outputs, likelihood) -- reference entropy_models.py:646-661.  Eval: outputs = round(x-mu)+mu.
        Training: outputs = x + noise (drawn on the device unless `noise` is injected; `mask` multiplies the
        draw as in the reference) and the likelihood is evaluated at outputs."""
        training = self.training if training is None else training
        if training and noise is None:
            noise = self._draw_noise(inputs, mask)
        out, lik, _, _ = self._run(inputs, scales, means, (), noise=noise if training else None)
        return out, lik

    @torch.no_grad()
    def quantize(self, inputs, mode, means=None, mask=None, noise=None):
        """reference entropy_models.py:126-152."""
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            noise = self._draw_noise(inputs, mask) if noise is None else noise
            out, _, _, _ = self._run(inputs, torch.ones_like(inputs), means, (), noise=noise)
            return out
        out, _, sym, _ = self._run(inputs, torch.ones_like(inputs), means, ("sym",))
        return out if mode == "dequantize" else sym

    @torch.no_grad()
    def compress(self, inputs, indexes, means=None):
        """reference entropy_models.py:205-238."""
        return self._encode_symbols(self.quantize(inputs, "symbols", means), indexes)

    @torch.no_grad()
    def decompress(self, strings, indexes, means=None):
        """reference entropy_models.py:241-287."""
        sym = self._decode_symbols(strings, indexes)
        return self.dequantize(sym.to(indexes.device) if means is None else sym.to(means.device), means)

    @torch.no_grad()
    def build_indexes(self, scales):
        """reference entropy_models.py:663-668, one binary search instead of 63 compare passes."""
        _, _, _, idx = self._run(torch.zeros_like(scales), scales, None, ("idx",))
        return idx
