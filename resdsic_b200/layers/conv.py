"""Convolution-family layers on the implicit-GEMM kernel.

Mirrors of nn.Conv2d / nn.ConvTranspose2d / nn.Linear / nn.PixelShuffle as the
reference uses them (WACNN/utils.py:116-134, layers/layers.py:29-43,
layers/win_attention.py:76-78): same parameter names, shapes and default init.
"""
import math

import torch
import torch.nn as nn

from .. import _lib, packing
from .base import B200Module, Ctx


def _default_conv_init(weight, bias):
    """The reference's EFFECTIVE init.  `CompressionModel.__init__` calls `_initialize_weights()` (the
    Kaiming-normal / zero-bias loop, WACNN/base.py:19-20,29-34) from `super().__init__()` at cnn.py:26-27,
    i.e. BEFORE any sub-module exists, so that loop visits nothing and every nn.Conv2d / nn.ConvTranspose2d
    keeps torch's `reset_parameters()` default: kaiming_uniform_(a=sqrt(5)) weights and U(+-1/sqrt(fan_in))
    biases.  Same calls in the same order as torch's `_ConvNd.reset_parameters`, so that under one
    `torch.manual_seed` this constructor draws the reference constructor's weights bit for bit
    (tests/test_refinit.py)."""
    nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    fan_in, _ = nn.init._calculate_fan_in_and_fan_out(weight)
    if fan_in != 0:
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(bias, -bound, bound)


class Conv2d(B200Module):
    """nn.Conv2d(in, out, k, stride, padding=k//2) with bias."""

    fuses_gelu = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.padding = kernel_size // 2 if padding is None else padding
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        _default_conv_init(self.weight, self.bias)

    def packed(self, wdt):
        return self._packed(("w", wdt), (self.weight, self.bias), lambda: (
            packing.pack_conv_weight(self.weight, wdt), self.bias.detach().float().contiguous()))

    def can_fuse_gdn(self, ctx, x, gdn):
        """Fuse the following GDN into this conv's kernel when the conv is short-K (memory-bound): the fused
        kernel gives up pipeline stages for the resident gamma', which only pays off for cheap producers."""
        C = self.out_channels
        k_iters = self.kernel_size ** 2 * -(-self.in_channels // 64)
        patch = ctx.wdt_for(x) != torch.bfloat16 and self.kernel_size ** 2 * self.in_channels <= 256
        return C == gdn.in_channels and C % 64 == 0 and C <= 192 and (patch or (ctx.wdt_for(x) == torch.bfloat16 and k_iters <= 32))

    def emit(self, ctx: Ctx, x, gelu=False, epilogue=None, out=None, out_dtype=None, pixel_shuffle=0, want_sq=False,
             tail=None, **kw):
        k, s, p = self.kernel_size, self.stride, self.padding
        OH, OW = (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1
        if ctx.precision == "bf16" and ctx.wdt_for(x) != torch.bfloat16 and k * k * self.in_channels <= 256:
            # narrow-channel input (the RGB image): im2col into a bf16 patch tensor, then a pointwise
            # tensor-core GEMM with K = ceil16(k*k*Cin) -- instead of the fp32 SIMT kernel
            Kp = -(-(k * k * self.in_channels) // 16) * 16
            w, b = self._packed(("patch", Kp), (self.weight, self.bias), lambda: (
                packing.pack_conv_weight(self.weight, torch.bfloat16, k_pad_to=16), self.bias.detach().float().contiguous()))
            patches = ctx.prog.patchify(x, ctx.buf(x.B, OH, OW, Kp, torch.bfloat16), k, k, s, p)
            x, k, s, p = patches, 1, 1, 0
        else:
            w, b = self.packed(ctx.wdt_for(x))
        if pixel_shuffle == 2 and ctx.wdt_for(x) == torch.bfloat16 and (self.out_channels // 4) % 16 == 0 and not kw \
                and (out is None or (out.ld % 16 == 0 and out.coff % 16 == 0)):
            # PixelShuffle with the weight rows packed sub-position-major (column s * C/4 + c instead of 4 c + s): 16
            # consecutive GEMM columns are then 16 channels of one shuffled pixel -- the vector epilogue applies
            # (the scalar path cost 150 us for h_s's 256 -> 4 x 288 layer against 35 us of tensor time)
            Cq = self.out_channels // 4
            w, b = self._packed(("ps3", ctx.wdt_for(x)), (self.weight, self.bias), lambda: (
                packing.pack_conv_weight(self.weight.view(Cq, 4, *self.weight.shape[1:]).transpose(0, 1).reshape(self.weight.shape),
                                         ctx.wdt_for(x)),
                self.bias.detach().float().view(Cq, 4).t().reshape(-1).contiguous()))
            pixel_shuffle = 3
        if tail is not None:
            kw["tail"] = tail
        if out is None:
            if pixel_shuffle:
                out = ctx.buf(x.B, OH * 2, OW * 2, self.out_channels // 4, out_dtype)
            else:
                out = ctx.buf(x.B, OH, OW, tail[2] if tail is not None else self.out_channels, out_dtype)
        if epilogue is None:
            epilogue = _lib.EPI_GELU if gelu else _lib.EPI_NONE
        if want_sq:
            kw["out2"] = ctx.buf(out.B, out.H, out.W, out.C)
        y = ctx.prog.conv(x, w, b, self.out_channels, k, k, s, p, p, out, epilogue=epilogue,
                          pixel_shuffle=pixel_shuffle, OH=OH, OW=OW, **kw)
        return (y, kw["out2"]) if want_sq else y


def _emit_partial(self, ctx, x, part, split, out=None, res=None, gelu=False, out_dtype=None):
    """Split-K form of a conv over a channel concatenation cat(a[split], b): part 0 convolves `a` with
    W[:, :split] (no bias) into fp32 partial sums; part 1 convolves `b` with W[:, split:], adds the bias and
    the partial sums `res` (and GELU).  conv(cat(a,b); W) == part0 + part1, which lets the slice loop
    (cnn.py:165-182) pre-compute the latent-only part of every context transform off the serial chain."""
    wdt = ctx.wdt_for(x)
    w_lat, w_ext, b = self._packed(("split", wdt, split), (self.weight, self.bias), lambda: (
        packing.pack_conv_weight(self.weight[:, :split], wdt), packing.pack_conv_weight(self.weight[:, split:], wdt),
        self.bias.detach().float().contiguous()))
    k, s, p = self.kernel_size, self.stride, self.padding
    OH, OW = (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1
    if part == 0:
        assert x.C == split
        out = out if out is not None else ctx.buf(x.B, OH, OW, self.out_channels, torch.float32)
        return ctx.prog.conv(x, w_lat, None, self.out_channels, k, k, s, p, p, out, OH=OH, OW=OW)
    assert x.C == self.in_channels - split and res is not None
    out = out if out is not None else ctx.buf(x.B, OH, OW, self.out_channels, out_dtype)
    epi = _lib.EPI_RES_GELU if gelu else _lib.EPI_ADD_RES
    return ctx.prog.conv(x, w_ext, b, self.out_channels, k, k, s, p, p, out, epilogue=epi, res=res, OH=OH, OW=OW)


Conv2d.emit_partial = _emit_partial


def _packed_cols(conv, wdt, c0, c1):
    """Packed weight of `conv` restricted to input channels [c0, c1) (cached per conv and range)."""
    return conv._packed(("cols", wdt, c0, c1), (conv.weight, conv.bias), lambda: (
        packing.pack_conv_weight(conv.weight[:, c0:c1], wdt), conv.bias.detach().float().contiguous()))


def emit_grouped(ctx, owner, tag, convs, x, in_group_stride, cols=None, bias=True, **kw):
    """ONE launch for several convolutions of identical shape (the grouped form of rdsic_conv_desc): `convs[g]`
    convolves the channels `x.C` wide that start `g * in_group_stride` channels after `x`'s first (0: all share `x`)
    and produces output channels [g * Cout, (g + 1) * Cout) of the result.  `cols = (c0, c1)` restricts every conv to
    the input-channel range [c0, c1) of its weight (the split-K forms of the slice loop, cf. `Conv2d.emit_partial`);
    `bias=False` leaves the bias to another part of the split.  Remaining keywords (epilogue, res, out, out2 ...) go
    to `Program.conv`.  The concatenated weight / bias are cached on `owner` under `tag` and rebuilt when any member's
    packed tensors change."""
    c = convs[0]
    wdt = ctx.wdt_for(x)
    c0, c1 = cols if cols is not None else (0, c.in_channels)
    assert x.C == c1 - c0 and all(m.weight.shape == c.weight.shape for m in convs), (x.C, c0, c1)
    parts = [_packed_cols(m, wdt, c0, c1) for m in convs]
    key = tuple((w.data_ptr(), b.data_ptr()) for w, b in parts)
    cache = owner.__dict__.setdefault("_group_pack_cache", {})
    hit = cache.get((tag, wdt, c0, c1))
    if hit is None or hit[0] != key:
        with torch.no_grad():
            hit = (key, (torch.cat([w[:c.out_channels] for w, _ in parts], 0).contiguous(),
                         torch.cat([b for _, b in parts], 0).contiguous()), parts)  # (parts kept alive: the key holds their addresses)
        cache[(tag, wdt, c0, c1)] = hit
    w, b = hit[1]
    G, k, s_, p = len(convs), c.kernel_size, c.stride, c.padding
    OH, OW = (x.H + 2 * p - k) // s_ + 1, (x.W + 2 * p - k) // s_ + 1
    out = kw.pop("out", None)
    out_dtype = kw.pop("out_dtype", None)
    if out is None:
        out = ctx.buf(x.B, OH, OW, G * c.out_channels, out_dtype)
    if kw.pop("gelu", False):
        kw["epilogue"] = _lib.EPI_RES_GELU if kw.get("res") is not None else _lib.EPI_GELU
    elif "epilogue" not in kw and kw.get("res") is not None:
        kw["epilogue"] = _lib.EPI_ADD_RES
    return ctx.prog.conv(x, w, b if bias else None, G * c.out_channels, k, k, s_, p, p, out, OH=OH, OW=OW,
                         groups=G, in_group_stride=in_group_stride, **kw)


class ConvTranspose2d(B200Module):
    """nn.ConvTranspose2d(in, out, 5, stride=2, padding=2, output_padding=1): four
    sub-pixel phase GEMMs (3x3, 3x2, 2x3, 2x2 taps), output exactly 2x the input."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=2):
        super().__init__()
        if kernel_size != 5 or stride != 2:
            raise ValueError("only the reference's deconv(k=5, s=2) is supported")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels, 5, 5))
        self.bias = nn.Parameter(torch.empty(out_channels))
        _default_conv_init(self.weight, self.bias)

    def packed(self, wdt):
        return self._packed(("w", wdt), (self.weight, self.bias), lambda: (
            packing.pack_deconv_weight(self.weight, wdt), self.bias.detach().float().contiguous()))

    def can_fuse_gdn(self, ctx, x, gdn):
        C = self.out_channels
        return ctx.wdt_for(x) == torch.bfloat16 and C == gdn.in_channels and C % 64 == 0 and C <= 192 and \
            9 * -(-self.in_channels // 64) <= 32

    def emit(self, ctx: Ctx, x, out=None, out_dtype=None, want_sq=False, **kw):
        wdt = ctx.wdt_for(x)
        if wdt == torch.bfloat16 and self.out_channels * 4 <= 16 and not want_sq and not kw:
            # narrow head (192 -> 3): one merged 3x3 GEMM with 4*Cout columns + PixelShuffle addressing
            w, b4 = self._packed(("merged", wdt), (self.weight, self.bias), lambda: (
                packing.pack_deconv_merged(self.weight, wdt),
                self.bias.detach().float().repeat_interleave(4).contiguous()))
            if out is None:
                out = ctx.buf(x.B, 2 * x.H, 2 * x.W, self.out_channels, out_dtype)
            return ctx.prog.conv(x, w, b4, self.out_channels * 4, 3, 3, 1, 1, 1, out, pixel_shuffle=2, OH=x.H, OW=x.W)
        phases, b = self.packed(wdt)
        if out is None:
            out = ctx.buf(x.B, 2 * x.H, 2 * x.W, self.out_channels, out_dtype)
        if want_sq:
            kw["out2"] = ctx.buf(out.B, out.H, out.W, out.C)
        for (py, px), (w, R, S, ph, pw) in phases.items():
            ctx.prog.conv(x, w, b, self.out_channels, R, S, 1, ph, pw, out, OH=x.H, OW=x.W,
                          osy=2, osx=2, ooy=py, oox=px, **kw)
        return (out, kw["out2"]) if want_sq else out


class Linear(B200Module):
    """nn.Linear over the channel dim of a channels-last view (a 1x1 GEMM)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))  # nn.Linear default
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
            bound = 1 / math.sqrt(in_features)
            nn.init.uniform_(self.bias, -bound, bound)
        else:
            self.register_parameter("bias", None)

    def packed(self, wdt):
        tensors = (self.weight,) if self.bias is None else (self.weight, self.bias)
        return self._packed(("w", wdt), tensors, lambda: (
            packing.pack_linear_weight(self.weight, wdt),
            None if self.bias is None else self.bias.detach().float().contiguous()))

    def emit(self, ctx: Ctx, x, out=None, epilogue=_lib.EPI_NONE, out_dtype=None, **kw):
        w, b = self.packed(ctx.wdt_for(x))
        if out is None:
            out = ctx.buf(x.B, x.H, x.W, self.out_features, out_dtype)
        return ctx.prog.conv(x, w, b, self.out_features, 1, 1, 1, 0, 0, out, epilogue=epilogue, **kw)


class PixelShuffle(B200Module):
    """Parameter-less placeholder keeping `subpel_conv3x3`'s child index (".2.0.weight");
    the shuffle itself is folded into the producing conv's store addressing."""

    def __init__(self, r):
        super().__init__()
        self.r = r

    def emit(self, ctx, x, **kw):
        raise RuntimeError("PixelShuffle is fused into SubpelConv; it is never emitted on its own")


class SubpelConv(nn.Sequential, B200Module):
    """subpel_conv3x3 (reference layers/layers.py:34-38): conv3x3 to 4*C + PixelShuffle(2)."""

    fuses_gelu = True

    def __init__(self, in_ch, out_ch, r=2, kernel_size=3):
        if r != 2:
            raise ValueError("only r=2 is used by the reference")
        super().__init__(Conv2d(in_ch, out_ch * r * r, kernel_size), PixelShuffle(r))

    def emit(self, ctx, x, gelu=False, **kw):
        return self[0].emit(ctx, x, gelu=gelu, pixel_shuffle=2, **kw)

    forward = B200Module.forward


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """reference WACNN/utils.py:116-123"""
    return Conv2d(in_channels, out_channels, kernel_size, stride)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    """reference WACNN/utils.py:126-134"""
    return ConvTranspose2d(in_channels, out_channels, kernel_size, stride)


def conv3x3(in_ch, out_ch, stride=1):
    return Conv2d(in_ch, out_ch, 3, stride)


def conv1x1(in_ch, out_ch, stride=1):
    return Conv2d(in_ch, out_ch, 1, stride)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return SubpelConv(in_ch, out_ch, r)
