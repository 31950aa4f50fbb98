"""torch.autograd.Function nodes of the rate-distortion training step: every forward AND every backward is one of the
library's CUDA kernels, called through the C ABI (include/resdsic_b200.h, section "Training").  torch supplies the
graph walk, gradient accumulation and the optimisers -- no arithmetic of its own on activation-sized tensors.

Activations are fp32 channels-last tensors [B,H,W,C] (contiguous); likelihoods are NCHW like the module outputs.
Reference being replaced: torch autograd over `WACNN.forward` in train() mode (models/WACNN/cnn.py:143-193) with
the custom gradients of ops/bound_ops.py:21-27, ops/ops.py:34 and entropy_models.py:429-430.
"""
import ctypes as C

import torch

from .. import _lib, packing
from .._lib import AttnDesc, ConvDesc, EBDesc, GCDesc, View

_NULL = View()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check(t):
    if not t.is_cuda:
        raise RuntimeError("resdsic_b200 training runs on CUDA devices only (sm_100a kernels; no CPU fallback)")
    assert t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.is_contiguous())
    return t


def _view(t, ld, coff=0, nchw=False):
    v = View()
    v.ptr, v.dtype, v.ld, v.coff, v.nchw = t.data_ptr(), _lib.F32, ld, coff, int(nchw)
    return v


LAUNCHES = [0]  # kernels launched through this module (bench.py's `gpu_launches` of the training workload)


def _call(name, t, *args):
    LAUNCHES[0] += 1
    with torch.cuda.device(t.device):
        rc = getattr(_lib.lib(), name)(*args, _stream(t))
    _lib.check(rc, name)


def _conv_desc(x, B, H, W, Cin, w, b, Cout, KH, KW, stride, ph, pw, out, OH, OW, OHt=None, OWt=None, osy=1, osx=1, ooy=0, oox=0):
    d = ConvDesc()
    d.in_ = _view(x, Cin)
    d.B, d.H, d.W, d.Cin = B, H, W, Cin
    d.weight = w.data_ptr() if w is not None else None
    d.bias = b.data_ptr() if b is not None else None
    d.w_dtype = _lib.F32
    d.Cout, d.KH, d.KW, d.stride, d.pad_h, d.pad_w = Cout, KH, KW, stride, ph, pw
    d.OH, d.OW = OH, OW
    d.OHt, d.OWt = (OH if OHt is None else OHt), (OW if OWt is None else OWt)
    d.osy, d.osx, d.ooy, d.oox = osy, osx, ooy, oox
    d.epilogue = _lib.EPI_NONE
    d.out = _view(out, Cout)
    return d


# ----------------------------------------------------------------------------- pointwise
def _pw(op, a, b=None, c=None, n_out=1, alpha=0.0):
    a = _check(a)
    outs = [torch.empty_like(a) for _ in range(n_out)]
    p = lambda t: None if t is None else _check(t).data_ptr()
    _call("rdsic_pointwise_f32", a, op, a.numel(), a.data_ptr(), p(b), p(c), outs[0].data_ptr(),
          outs[1].data_ptr() if n_out > 1 else None, C.c_float(alpha))
    return outs[0] if n_out == 1 else outs


class GeluFn(torch.autograd.Function):
    """nn.GELU (exact erf form)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _pw(_lib.PW_GELU_FWD, x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return _pw(_lib.PW_GELU_BWD, g.contiguous(), x)


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return _pw(_lib.PW_ADD, a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


class GateFn(torch.autograd.Function):
    """out = a * sigmoid(b) + x  (Win_noShift_Attention.forward, layers/layers.py:83-89)."""

    @staticmethod
    def forward(ctx, a, b, x):
        ctx.save_for_backward(a, b)
        return _pw(_lib.PW_GATE_FWD, a, b, x)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        da, db = _pw(_lib.PW_GATE_BWD, g, a, b, n_out=2)
        return da, db, g


class SquareFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _pw(_lib.PW_SQUARE_FWD, x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return _pw(_lib.PW_SQUARE_BWD, g.contiguous(), x)


class GdnScaleFn(torch.autograd.Function):
    """y = x * rsqrt(norm)  (inverse: x * sqrt(norm))  -- layers/gdn.py:70-75."""

    @staticmethod
    def forward(ctx, x, norm, inverse):
        ctx.save_for_backward(x, norm)
        ctx.inverse = bool(inverse)
        return _pw(_lib.PW_GDN_FWD, x, norm, alpha=1.0 if inverse else 0.0)

    @staticmethod
    def backward(ctx, g):
        x, norm = ctx.saved_tensors
        dx, dn = _pw(_lib.PW_GDN_BWD, g.contiguous(), x, norm, n_out=2, alpha=1.0 if ctx.inverse else 0.0)
        return dx, dn, None


class LrpFn(torch.autograd.Function):
    """y_hat + 0.5 * tanh(lrp)  (cnn.py:179-182)."""

    @staticmethod
    def forward(ctx, y_hat, v):
        ctx.save_for_backward(v)
        return _pw(_lib.PW_LRP_FWD, y_hat, v)

    @staticmethod
    def backward(ctx, g):
        (v,) = ctx.saved_tensors
        g = g.contiguous()
        return g, _pw(_lib.PW_LRP_BWD, g, v)


class PixelShuffleFn(torch.autograd.Function):
    """nn.PixelShuffle(2) on channels-last data: [B,H,W,4C] -> [B,2H,2W,C]."""

    @staticmethod
    def forward(ctx, x):
        x = _check(x)
        B, H, W, C4 = x.shape
        out = torch.empty(B, 2 * H, 2 * W, C4 // 4, device=x.device, dtype=torch.float32)
        _call("rdsic_pixel_shuffle_f32", x, 0, x.data_ptr(), out.data_ptr(), B, H, W, C4 // 4)
        return out

    @staticmethod
    def backward(ctx, g):
        g = _check(g.contiguous())
        B, H2, W2, Cc = g.shape
        out = torch.empty(B, H2 // 2, W2 // 2, 4 * Cc, device=g.device, dtype=torch.float32)
        _call("rdsic_pixel_shuffle_f32", g, 1, g.data_ptr(), out.data_ptr(), B, H2 // 2, W2 // 2, Cc)
        return out


# ----------------------------------------------------------------------------- layout / concatenation (copy kernel)
def _copy(src_tv, dst_tv):
    from ..program import Program
    prog = Program(src_tv.t.device)
    prog.copy(src_tv, dst_tv)
    prog.run()


class ToChannelsLastFn(torch.autograd.Function):
    """NCHW -> NHWC (the library's copy kernel; the reference's modules work on NCHW throughout)."""

    @staticmethod
    def forward(ctx, x):
        from ..program import TV
        x = _check(x.contiguous())
        B, Cc, H, W = x.shape
        out = torch.empty(B, H, W, Cc, device=x.device, dtype=torch.float32)
        _copy(TV.nchw_of(x), TV(out, B, H, W, Cc))
        return out

    @staticmethod
    def backward(ctx, g):
        return ToChannelsFirstFn.apply(g)


class ToChannelsFirstFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from ..program import TV
        x = _check(x.contiguous())
        B, H, W, Cc = x.shape
        out = torch.empty(B, Cc, H, W, device=x.device, dtype=torch.float32)
        _copy(TV(x, B, H, W, Cc), TV.nchw_of(out))
        return out

    @staticmethod
    def backward(ctx, g):
        return ToChannelsLastFn.apply(g)


class CatChannelsFn(torch.autograd.Function):
    """torch.cat(dim=channels) of channels-last tensors (the slice loop's support concatenations, cnn.py:165-178)."""

    @staticmethod
    def forward(ctx, *xs):
        from ..program import TV, Program
        B, H, W = xs[0].shape[:3]
        ctx.sizes = [x.shape[3] for x in xs]
        Ct = sum(ctx.sizes)
        out = torch.empty(B, H, W, Ct, device=xs[0].device, dtype=torch.float32)
        prog, off = Program(out.device), 0
        for x in xs:
            x = _check(x.contiguous())
            prog.copy(TV(x, B, H, W, x.shape[3]), TV(out, B, H, W, x.shape[3], ld=Ct, coff=off))
            off += x.shape[3]
        prog.run()
        return out

    @staticmethod
    def backward(ctx, g):
        from ..program import TV, Program
        g = _check(g.contiguous())
        B, H, W, Ct = g.shape
        prog, off, outs = Program(g.device), 0, []
        for n in ctx.sizes:
            o = torch.empty(B, H, W, n, device=g.device, dtype=torch.float32)
            prog.copy(TV(g, B, H, W, n, ld=Ct, coff=off), TV(o, B, H, W, n))
            outs.append(o)
            off += n
        prog.run()
        return tuple(outs)


class CatNchwFn(torch.autograd.Function):
    """torch.cat(dim=1) of NCHW tensors (the ten likelihood slices, cnn.py:186)."""

    @staticmethod
    def forward(ctx, *xs):
        from ..program import TV, Program
        B, _, H, W = xs[0].shape
        ctx.sizes = [x.shape[1] for x in xs]
        out = torch.empty(B, sum(ctx.sizes), H, W, device=xs[0].device, dtype=torch.float32)
        prog, off = Program(out.device), 0
        for x in xs:
            x = _check(x.contiguous())
            prog.copy(TV.nchw_of(x), TV.nchw_channels(out, off, x.shape[1]))
            off += x.shape[1]
        prog.run()
        return out

    @staticmethod
    def backward(ctx, g):
        from ..program import TV, Program
        g = _check(g.contiguous())
        B, _, H, W = g.shape
        prog, off, outs = Program(g.device), 0, []
        for n in ctx.sizes:
            o = torch.empty(B, n, H, W, device=g.device, dtype=torch.float32)
            prog.copy(TV.nchw_channels(g, off, n), TV.nchw_of(o))
            outs.append(o)
            off += n
        prog.run()
        return tuple(outs)


class SliceChannelsFn(torch.autograd.Function):
    """x[..., off:off+n] as a contiguous tensor (y.chunk(10, 1), cnn.py:157)."""

    @staticmethod
    def forward(ctx, x, off, n):
        from ..program import TV
        x = _check(x.contiguous())
        B, H, W, Ct = x.shape
        ctx.meta = (off, n, Ct)
        out = torch.empty(B, H, W, n, device=x.device, dtype=torch.float32)
        _copy(TV(x, B, H, W, n, ld=Ct, coff=off), TV(out, B, H, W, n))
        return out

    @staticmethod
    def backward(ctx, g):
        from ..program import TV
        off, n, Ct = ctx.meta
        g = _check(g.contiguous())
        B, H, W, _ = g.shape
        out = torch.zeros(B, H, W, Ct, device=g.device, dtype=torch.float32)  # (cudaMemset: no arithmetic)
        _copy(TV(g, B, H, W, n), TV(out, B, H, W, n, ld=Ct, coff=off))
        return out, None, None


# ----------------------------------------------------------------------------- convolutions
class ConvFn(torch.autograd.Function):
    """nn.Conv2d / nn.Linear (1x1) on channels-last data.  forward: rdsic_conv_forward (fp32 implicit GEMM);
    backward: rdsic_conv_dgrad_f32 + rdsic_conv_wgrad_f32.  `weight` is the module's OIHW parameter."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding):
        x = _check(x.contiguous())
        B, H, W, Cin = x.shape
        Cout, _, KH, KW = weight.shape
        OH, OW = (H + 2 * padding - KH) // stride + 1, (W + 2 * padding - KW) // stride + 1
        wp = packing.pack_conv_weight(weight, torch.float32)
        b = None if bias is None else bias.detach().float().contiguous()
        out = torch.empty(B, OH, OW, Cout, device=x.device, dtype=torch.float32)
        d = _conv_desc(x, B, H, W, Cin, wp, b, Cout, KH, KW, stride, padding, padding, out, OH, OW)
        _call("rdsic_conv_forward", x, C.byref(d))
        ctx.save_for_backward(x, weight)
        ctx.geom = (stride, padding, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        stride, padding, has_bias = ctx.geom
        g = _check(g.contiguous())
        B, H, W, Cin = x.shape
        Cout, _, KH, KW = weight.shape
        OH, OW = g.shape[1], g.shape[2]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # forward weight as [Cin][KH*KW*Cout] (tap-major, output-channel-minor)
            wd = weight.detach().permute(1, 2, 3, 0).reshape(Cin, KH * KW * Cout).contiguous()
            dx = torch.empty_like(x)
            d = _conv_desc(dx, B, H, W, Cin, None, None, Cout, KH, KW, stride, padding, padding, g, OH, OW)
            _call("rdsic_conv_dgrad_f32", x, C.byref(d), wd.data_ptr())
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            dwp = torch.empty(Cout, KH * KW * Cin, device=x.device, dtype=torch.float32)
            dbt = torch.empty(Cout, device=x.device, dtype=torch.float32) if has_bias else None
            d = _conv_desc(x, B, H, W, Cin, None, None, Cout, KH, KW, stride, padding, padding, g, OH, OW)
            _call("rdsic_conv_wgrad_f32", x, C.byref(d), dwp.data_ptr(), None if dbt is None else dbt.data_ptr())
            dw = dwp.view(Cout, KH, KW, Cin).permute(0, 3, 1, 2)  # packed [Cout][r][s][ci] -> OIHW (a view)
            db = dbt
        return dx, dw, db, None, None


class DeconvFn(torch.autograd.Function):
    """nn.ConvTranspose2d(k=5, s=2, p=2, output_padding=1) (WACNN/utils.py:126-134).  forward: four sub-pixel phase
    GEMMs; d input: the stride-2 5x5 convolution of d out with the same weight tensor (rdsic_conv_forward);
    d weight: rdsic_conv_wgrad_f32 with the roles of x and d out exchanged."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _check(x.contiguous())
        B, H, W, Cin = x.shape
        Cout = weight.shape[1]
        phases = packing.pack_deconv_weight(weight, torch.float32)
        b = bias.detach().float().contiguous()
        out = torch.empty(B, 2 * H, 2 * W, Cout, device=x.device, dtype=torch.float32)
        for (py, px), (w, R, S, ph, pw) in phases.items():
            d = _conv_desc(x, B, H, W, Cin, w, b, Cout, R, S, 1, ph, pw, out, H, W, 2 * H, 2 * W, 2, 2, py, px)
            _call("rdsic_conv_forward", x, C.byref(d))
        ctx.save_for_backward(x, weight)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = _check(g.contiguous())
        B, H, W, Cin = x.shape
        Cout = weight.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wc = packing.pack_conv_weight(weight, torch.float32)  # [Cin(out)][25 * Cout(in)]
            dx = torch.empty_like(x)
            d = _conv_desc(g, B, 2 * H, 2 * W, Cout, wc, None, Cin, 5, 5, 2, 2, 2, dx, H, W)
            _call("rdsic_conv_forward", x, C.byref(d))
        if ctx.needs_input_grad[1]:
            dwp = torch.empty(Cin, 25 * Cout, device=x.device, dtype=torch.float32)
            d = _conv_desc(g, B, 2 * H, 2 * W, Cout, None, None, Cin, 5, 5, 2, 2, 2, x, H, W)
            _call("rdsic_conv_wgrad_f32", x, C.byref(d), dwp.data_ptr(), None)
            dw = dwp.view(Cin, 5, 5, Cout).permute(0, 3, 1, 2)  # -> [Cin, Cout, 5, 5]
        if ctx.needs_input_grad[2]:
            # d bias = column sums of d out: the d-bias output of a 1x1 weight-gradient call over d out itself
            scratch = torch.empty(Cout, Cout, device=x.device, dtype=torch.float32)
            db = torch.empty(Cout, device=x.device, dtype=torch.float32)
            d = _conv_desc(g, B, 2 * H, 2 * W, Cout, None, None, Cout, 1, 1, 1, 0, 0, g, 2 * H, 2 * W)
            _call("rdsic_conv_wgrad_f32", x, C.byref(d), scratch.data_ptr(), db.data_ptr())
        return dx, dw, db


# ----------------------------------------------------------------------------- window attention
class WindowAttentionFn(torch.autograd.Function):
    """roll + partition + softmax((q*scale) k^T + bias + mask) v + reverse + roll back (win_attention.py:94-112,
    159-200) on the qkv map [B,H,W,3C]."""

    @staticmethod
    def forward(ctx, qkv, bias_table, heads, ws, shift, scale):
        qkv = _check(qkv.contiguous())
        B, H, W, C3 = qkv.shape
        Cc = C3 // 3
        tab = bias_table.detach().float().contiguous()
        out = torch.empty(B, H, W, Cc, device=qkv.device, dtype=torch.float32)
        d = AttnDesc()
        d.qkv, d.out, d.bias_table = _view(qkv, C3), _view(out, Cc), tab.data_ptr()
        d.B, d.H, d.W, d.C, d.heads, d.ws, d.shift, d.scale = B, H, W, Cc, heads, ws, shift, scale
        _call("rdsic_attn_forward", qkv, C.byref(d))
        ctx.save_for_backward(qkv, tab)
        ctx.meta = (heads, ws, shift, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, tab = ctx.saved_tensors
        heads, ws, shift, scale = ctx.meta
        g = _check(g.contiguous())
        B, H, W, C3 = qkv.shape
        Cc = C3 // 3
        dqkv = torch.empty_like(qkv)
        dtab = torch.empty_like(tab)
        d = AttnDesc()
        d.qkv, d.out, d.bias_table = _view(qkv, C3), _NULL, tab.data_ptr()
        d.B, d.H, d.W, d.C, d.heads, d.ws, d.shift, d.scale = B, H, W, Cc, heads, ws, shift, scale
        gv, dv = _view(g, Cc), _view(dqkv, C3)
        _call("rdsic_attn_backward_f32", qkv, C.byref(d), C.byref(gv), C.byref(dv), dtab.data_ptr())
        return dqkv, dtab, None, None, None, None


# ----------------------------------------------------------------------------- entropy models
class GaussianConditionalFn(torch.autograd.Function):
    """GaussianConditional.forward in training mode + the slice loop's ste_round (entropy_models.py:646-661,
    cnn.py:175-177): (y, mu, scale, noise) -> (likelihood [B,Cs,h,w] NCHW, y_hat = round(y - mu) + mu)."""

    @staticmethod
    def forward(ctx, y, mu, scale, noise, table, scale_bound, lik_bound):
        y, mu, scale, noise = (_check(t.contiguous()) for t in (y, mu, scale, noise))
        B, h, w, Cs = y.shape
        lik = torch.empty(B, Cs, h, w, device=y.device, dtype=torch.float32)
        y_hat = torch.empty_like(y)
        d = GaussianConditionalFn._desc(y, mu, scale, noise, table, scale_bound, lik_bound, lik)
        d.y_hat[0] = _view(y_hat, Cs)
        _call("rdsic_gc_forward", y, C.byref(d))
        ctx.save_for_backward(y, mu, scale, noise, table)
        ctx.bounds = (scale_bound, lik_bound)
        return lik, y_hat

    @staticmethod
    def _desc(y, mu, scale, noise, table, scale_bound, lik_bound, lik):
        B, h, w, Cs = y.shape
        d = GCDesc()
        d.y, d.mu, d.scale, d.noise = _view(y, Cs), _view(mu, Cs), _view(scale, Cs), _view(noise, Cs)
        d.lik = lik.data_ptr() if lik is not None else None
        d.table, d.n_table = table.data_ptr(), table.numel()
        d.B, d.h, d.w, d.Cs, d.Ctot, d.lik_coff = B, h, w, Cs, Cs, 0
        d.scale_bound, d.lik_bound = scale_bound, lik_bound
        return d

    @staticmethod
    def backward(ctx, g_lik, g_yhat):
        y, mu, scale, noise, table = ctx.saved_tensors
        scale_bound, lik_bound = ctx.bounds
        Cs = y.shape[3]
        g_lik = _check(g_lik.contiguous()) if g_lik is not None else torch.zeros(y.shape[0], Cs, y.shape[1], y.shape[2], device=y.device)
        dy, dmu, dscale = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
        d = GaussianConditionalFn._desc(y, mu, scale, noise, table, scale_bound, lik_bound, None)
        gy = _view(_check(g_yhat.contiguous()), Cs) if g_yhat is not None else None
        vy, vm, vs = _view(dy, Cs), _view(dmu, Cs), _view(dscale, Cs)
        _call("rdsic_gc_backward", y, C.byref(d), g_lik.data_ptr(), C.byref(gy) if gy is not None else None,
              C.byref(vy), C.byref(vm), C.byref(vs))
        return dy, dmu, dscale, None, None, None, None


class EntropyBottleneckFn(torch.autograd.Function):
    """EntropyBottleneck.forward in training mode + the hyper path's ste_round (entropy_models.py:447-490,
    cnn.py:150-154): (z, packed parameters [C][60], noise) -> (likelihood [B,C,h,w], z_hat = round(z - med) + med).
    The gradient w.r.t. the packed parameters flows on through torch's softplus / tanh of the module parameters."""

    @staticmethod
    def forward(ctx, z, params, noise, lik_bound):
        z, noise = _check(z.contiguous()), _check(noise.contiguous())
        pk = _check(params.detach().contiguous())
        B, h, w, Cc = z.shape
        lik = torch.empty(B, Cc, h, w, device=z.device, dtype=torch.float32)
        z_hat = torch.empty_like(z)
        d = EntropyBottleneckFn._desc(z, pk, noise, lik_bound)
        d.z_hat, d.lik = _view(z_hat, Cc), lik.data_ptr()
        _call("rdsic_eb_forward", z, C.byref(d))
        ctx.save_for_backward(z, pk, noise)
        ctx.lik_bound = lik_bound
        return lik, z_hat

    @staticmethod
    def _desc(z, pk, noise, lik_bound):
        B, h, w, Cc = z.shape
        d = EBDesc()
        d.z, d.noise, d.params = _view(z, Cc), _view(noise, Cc), pk.data_ptr()
        d.B, d.h, d.w, d.C, d.lik_bound = B, h, w, Cc, lik_bound
        return d

    @staticmethod
    def backward(ctx, g_lik, g_zhat):
        z, pk, noise = ctx.saved_tensors
        Cc = z.shape[3]
        g_lik = _check(g_lik.contiguous())
        dz, dpk = torch.empty_like(z), torch.empty_like(pk)
        d = EntropyBottleneckFn._desc(z, pk, noise, ctx.lik_bound)
        gz = _view(_check(g_zhat.contiguous()), Cc) if g_zhat is not None else None
        vz = _view(dz, Cc)
        _call("rdsic_eb_backward", z, C.byref(d), g_lik.data_ptr(), C.byref(gz) if gz is not None else None, C.byref(vz),
              dpk.data_ptr())
        return dz, dpk, None, None


class AuxLossFn(torch.autograd.Function):
    """EntropyBottleneck.loss (entropy_models.py:396-399): sum |logits(quantiles) - target|, gradient to `quantiles`
    only (every other parameter is detached there)."""

    @staticmethod
    def forward(ctx, quantiles, packed, target):
        q = _check(quantiles.detach().contiguous())
        pk, tg = _check(packed.detach().contiguous()), _check(target.detach().contiguous())
        Cc = q.shape[0]
        out = torch.empty(1, device=q.device, dtype=torch.float32)
        with torch.cuda.device(q.device):
            rc = _lib.lib().rdsic_eb_aux_loss(pk.data_ptr(), q.data_ptr(), tg.data_ptr(), Cc, None, out.data_ptr(), _stream(q))
        _lib.check(rc, "rdsic_eb_aux_loss")
        ctx.save_for_backward(q, pk, tg)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        q, pk, tg = ctx.saved_tensors
        dq = torch.empty_like(q)
        _call("rdsic_eb_aux_backward", q, pk.data_ptr(), q.data_ptr(), tg.data_ptr(), q.shape[0], C.c_float(float(g)), dq.data_ptr())
        return dq, None, None


# ----------------------------------------------------------------------------- loss reductions
def _reduce(op, a, b=None):
    a = _check(a.contiguous())
    out = torch.empty(1, device=a.device, dtype=torch.float64)
    _call("rdsic_reduce_f32", a, op, a.numel(), a.data_ptr(), None if b is None else _check(b.contiguous()).data_ptr(), out.data_ptr())
    return out


class SumLogFn(torch.autograd.Function):
    """sum(log(likelihoods)) (training/loss.py:24), fp64 accumulation."""

    @staticmethod
    def forward(ctx, lik):
        ctx.save_for_backward(lik)
        return _reduce(_lib.RED_SUM_LOG, lik)[0].float()

    @staticmethod
    def backward(ctx, g):
        (lik,) = ctx.saved_tensors
        return _pw(_lib.PW_RECIP_SCALE, lik.contiguous(), alpha=float(g))


class MseFn(torch.autograd.Function):
    """nn.MSELoss()(x_hat, target) (training/loss.py:27)."""

    @staticmethod
    def forward(ctx, x_hat, target):
        ctx.save_for_backward(x_hat, target)
        return (_reduce(_lib.RED_SSE, x_hat, target)[0] / x_hat.numel()).float()

    @staticmethod
    def backward(ctx, g):
        x_hat, target = ctx.saved_tensors
        return _pw(_lib.PW_DIFF_SCALE, x_hat.contiguous(), target.contiguous(), alpha=2.0 * float(g) / x_hat.numel()), None
