"""Win_noShift_Attention + ResidualUnit (reference layers/layers.py:45-89).

Every GELU, the residual add, the sigmoid gate and the identity add are
epilogues of the producing GEMM: a ResidualUnit is 3 launches (the reference:
3 convs + 3 GELUs + 1 add), the gate `a*sigmoid(b) + x` rides on conv_b.4."""
import torch
import torch.nn as nn

from .. import _lib
from .base import GELU, B200Module, Ctx, Sequential
from .conv import conv1x1, conv3x3
from .win_attention import WinBasedAttention


class ResidualUnit(B200Module):
    fuse_tail = True  # bf16 mode: fuse conv3x3 + conv1x1 tail (class-level switch for A/B measurements)

    def __init__(self, N):
        super().__init__()
        self.conv = Sequential(conv1x1(N, N // 2), GELU(), conv3x3(N // 2, N // 2), GELU(), conv1x1(N // 2, N))
        self.relu = GELU()

    def emit(self, ctx: Ctx, x, **kw):
        t = self.conv[0].emit(ctx, x, gelu=True)
        c3, c1 = self.conv[2], self.conv[4]
        if (ctx.wdt_for(t) == torch.bfloat16 and ctx.wdt_for(x) == torch.bfloat16 and c1.out_channels <= 192
                and c3.out_channels + c3.out_channels // 2 + c1.out_channels <= 480 and ResidualUnit.fuse_tail):
            # 3x3 conv -> GELU -> 1x1 conv -> +x -> GELU in ONE kernel (the 3x3's result feeds the 1x1 GEMM
            # from tensor memory): one launch and one activation round trip less per unit
            w3, b3 = c1.packed(torch.bfloat16)
            return c3.emit(ctx, t, tail=(w3, b3, c1.out_channels), res=x)
        t = c3.emit(ctx, t, gelu=True)
        return c1.emit(ctx, t, epilogue=_lib.EPI_RES_GELU, res=x)


class Win_noShift_Attention(B200Module):
    """Window-based self-attention gate: out = conv_a(x) * sigmoid(conv_b(x)) + x."""

    def __init__(self, dim, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        N = dim
        self.conv_a = Sequential(ResidualUnit(N), ResidualUnit(N), ResidualUnit(N))
        self.conv_b = Sequential(
            WinBasedAttention(dim=dim, num_heads=num_heads, window_size=window_size, shift_size=shift_size),
            ResidualUnit(N), ResidualUnit(N), ResidualUnit(N), conv1x1(N, N))

    def emit(self, ctx: Ctx, x, out=None, out_dtype=None, **kw):
        side = 1 if ctx.prog._lane != 1 else 2
        ctx.prog.fork(side)  # trunk (conv_a) || attention branch (conv_b)
        with ctx.prog.side(side):
            a = self.conv_a.emit(ctx, x)
        b = x
        for m in list(self.conv_b)[:4]:
            b = m.emit(ctx, b)
        ctx.prog.join(side)
        return self.conv_b[4].emit(ctx, b, epilogue=_lib.EPI_GATE, aux=a, res=x, out=out, out_dtype=out_dtype, **kw)
