"""Common machinery of the B200 layer modules.

Every layer keeps the reference's parameter/buffer names (so reference
state_dicts load unchanged) and knows how to *emit* itself into a `Program`
(resdsic_b200/program.py) over channels-last views.  `forward(x)` is the
reference-facing call: NCHW tensor in, NCHW tensor out, executed by the CUDA
library -- there is no torch/ATen compute path.
"""
import torch
import torch.nn as nn

from ..program import TV, Program

PRECISIONS = ("fp32", "bf16")


class Ctx:
    """Build context: the program being assembled + activation/weight dtypes."""

    def __init__(self, device, precision="fp32", build_only=False):
        assert precision in PRECISIONS, precision
        device = torch.device(device)
        # build_only: assemble descriptors over CPU tensors so that tests can check the
        # host-side graph builder without a GPU (tests/program_sim.py); such a program
        # can never be run by the product (Program.run refuses non-CUDA devices).
        if device.type != "cuda" and not build_only:
            raise RuntimeError("resdsic_b200 runs on CUDA devices only (sm_100a kernels; no CPU fallback)")
        self.device = device
        self.precision = precision
        self.act = torch.float32 if precision == "fp32" else torch.bfloat16
        self.wdt = self.act
        self.prog = Program(device)

    def wdt_for(self, x):
        """Weight dtype for a GEMM reading view `x`: the tensor-core kernel needs bf16 channels-last
        activations with Cin % 16 == 0; anything else (the fp32 NCHW input image) takes the fp32 kernel."""
        if self.precision == "bf16" and x.t.dtype == torch.bfloat16 and not x.nchw and x.C % 16 == 0:
            return torch.bfloat16
        return torch.float32

    def buf(self, B, H, W, C, dtype=None, ld=None):
        return TV.empty(B, H, W, C, dtype or self.act, self.device, ld)

    def from_nchw(self, x, dtype=None):
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        B, C, H, W = x.shape
        out = self.buf(B, H, W, C, dtype)
        self.prog.keep.append(x)
        return self.prog.copy(TV.nchw_of(x), out)

    def to_nchw(self, tv):
        out = torch.empty(tv.B, tv.C, tv.H, tv.W, dtype=torch.float32, device=self.device)
        self.prog.copy(tv, TV.nchw_of(out))
        return out


def _tree_to(obj, dev):
    if torch.is_tensor(obj):
        return obj.to(dev)
    if isinstance(obj, (tuple, list)):
        return type(obj)(_tree_to(o, dev) for o in obj)
    if isinstance(obj, dict):
        return {k: _tree_to(v, dev) for k, v in obj.items()}
    return obj


class B200Module(nn.Module):
    """nn.Module whose compute is a resdsic_b200 program."""

    precision = "fp32"

    def set_precision(self, precision):
        assert precision in PRECISIONS, precision
        for m in self.modules():
            if isinstance(m, B200Module):
                m.precision = precision
        return self

    # Weight packing is one-time host-side preprocessing.  With `pack_on_host` (default) the layout transforms run on
    # CPU copies of the parameters and only the packed result is copied to the device, so that no ATen kernel is
    # launched on the GPU on behalf of this package (every GPU launch of an inference process is then one of the
    # library's own kernels).  Training re-packs after every optimiser step and sets it False (device-side packing).
    pack_on_host = True

    # -- packed-weight cache, invalidated when a parameter changes version/storage/device
    def _packed(self, tag, tensors, fn):
        key = (tag,) + tuple((t.data_ptr(), t._version, str(t.device), t.dtype) for t in tensors)
        cache = self.__dict__.setdefault("_pack_cache", {})
        hit = cache.get(tag)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                dev = tensors[0].device
                if B200Module.pack_on_host and dev.type == "cuda":
                    hit = (key, _tree_to(self._call_on_host(fn), dev))
                else:
                    hit = (key, fn())
            cache[tag] = hit
        return hit[1]

    def _call_on_host(self, fn):
        """Run `fn` with this module's own parameters / buffers temporarily replaced by CPU copies."""
        saved = []
        for store in (self._parameters, self._buffers):
            for name, t in list(store.items()):
                if t is not None and t.device.type == "cuda":
                    saved.append((store, name, t))
                    store[name] = t.detach().cpu()
        try:
            return fn()
        finally:
            for store, name, t in saved:
                store[name] = t

    def emit(self, ctx: Ctx, x: TV, **kw) -> TV:  # pragma: no cover - interface
        raise NotImplementedError

    @torch.no_grad()
    def forward(self, x):
        ctx = Ctx(x.device, self.precision)
        y = self.emit(ctx, ctx.from_nchw(x))
        out = ctx.to_nchw(y)
        ctx.prog.run()
        return out


class Sequential(nn.Sequential, B200Module):
    """nn.Sequential of B200 modules (keeps the reference's integer child names)."""

    def emit(self, ctx, x, last_kw=None, mods=None, **kw):
        """`last_kw`: extra emit() arguments for the final module (output placement/dtype).  `mods`: emit this
        sub-chain instead of all children (see emit_modules)."""
        mods = list(self) if mods is None else list(mods)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            step = 2 if (getattr(m, "fuses_gelu", False) and isinstance(nxt, GELU)) else 1
            extra = dict(last_kw or {}) if i + step >= len(mods) else {}
            if step == 2:  # fuse "conv -> GELU" pairs into the conv epilogue
                x = m.emit(ctx, x, gelu=True, **extra)
            elif (getattr(nxt, "wants_square", False) and ctx.precision == "bf16" and hasattr(m, "weight")
                  and getattr(m, "can_fuse_gdn", lambda *_: False)(ctx, x, nxt)):
                # conv/deconv -> GDN in ONE kernel (x stays in TMEM, gamma' resident in smem)
                extra = dict(last_kw or {}) if i + 2 >= len(mods) else {}
                x = m.emit(ctx, x, gdn=nxt.fused_params(), **extra)
                step = 2
            elif getattr(nxt, "wants_square", False) and ctx.precision == "bf16" and hasattr(m, "weight"):
                # conv/deconv -> GDN in bf16 mode: the producer also stores x^2 (the GDN GEMM's A operand)
                x = m.emit(ctx, x, out2_square=True, want_sq=True, **extra)
                x, x2 = x
                extra = dict(last_kw or {}) if i + 2 >= len(mods) else {}
                x = nxt.emit(ctx, x, x2=x2, **extra)
                step = 2
            else:
                x = m.emit(ctx, x, **extra)
            i += step
        return x

    forward = B200Module.forward


def emit_modules(ctx, mods, x, last_kw=None):
    """Emit a chain of modules with the conv->GELU / conv->GDN fusions of Sequential.emit -- used on sub-ranges of a
    Sequential (the scalable models split g_a after its sixth child, scalable/single_decoder.py:196-202)."""
    return Sequential.emit(None, ctx, x, last_kw=last_kw, mods=mods)


class GELU(B200Module):
    """nn.GELU() (exact erf form).  Normally fused into the preceding conv."""

    def emit(self, ctx, x, **kw):
        out = ctx.buf(x.B, x.H, x.W, x.C)
        return ctx.prog.copy(x, out, op_code=1)
