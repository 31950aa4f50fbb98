"""Shifted-window attention (reference layers/win_attention.py).

`WinBasedAttention.forward` there is: NCHW->NHWC permute, mask build, roll,
window_partition, qkv Linear, scaled QK^T + relative-position bias + mask,
softmax, @V, proj Linear, window_reverse, roll back, permute, + shortcut.
Here it is three launches on channels-last data:
  1. qkv 1x1 GEMM on the UNshifted map (a per-pixel Linear commutes with the
     roll / partition permutations),
  2. the fused window kernel (roll / partition / reverse folded into addressing,
     analytic shift mask),
  3. proj 1x1 GEMM with the shortcut add as its epilogue.
"""
import torch
import torch.nn as nn

from .. import _lib
from .base import B200Module, Ctx
from .conv import Linear


def relative_position_index(ws):
    """reference win_attention.py:64-74 (closed form)."""
    t = torch.arange(ws * ws)
    hi, wi = t // ws, t % ws
    return (hi[:, None] - hi[None, :] + ws - 1) * (2 * ws - 1) + (wi[:, None] - wi[None, :] + ws - 1)


class WindowAttention(B200Module):
    def __init__(self, dim=192, window_size=(8, 8), num_heads=8, qkv_bias=True, qk_scale=None,
                 attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if not qkv_bias:
            raise ValueError("the reference always uses qkv_bias=True")
        if attn_drop or proj_drop:
            raise ValueError("dropout is always 0 in the reference (win_attention.py:50,136-137)")
        if window_size[0] != window_size[1]:
            raise ValueError("square windows only")
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        ws = window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(ws))
        self.qkv = Linear(dim, dim * 3)
        self.proj = Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02, a=-2.0, b=2.0)

    def bias_table(self):
        t = self.relative_position_bias_table
        return self._packed("tab", (t,), lambda: t.detach().float().contiguous())

    def emit(self, ctx: Ctx, x, shift=0, shortcut=None, **kw):
        """x: [B,H,W,C] unshifted map; returns shortcut + proj(attn) (or proj(attn) if no shortcut)."""
        ws = self.window_size[0]
        if x.H % ws or x.W % ws:
            raise ValueError(f"feature map {x.H}x{x.W} is not a multiple of the window size {ws} "
                             "(pad the image to a multiple of 64, reference eval_model/__main__.py:89-101)")
        qkv = self.qkv.emit(ctx, x)
        att = ctx.buf(x.B, x.H, x.W, self.dim)
        ctx.prog.attn(qkv, att, self.bias_table(), self.num_heads, ws, shift, float(self.scale))
        if shortcut is not None:
            return self.proj.emit(ctx, att, epilogue=_lib.EPI_ADD_RES, res=shortcut)
        return self.proj.emit(ctx, att)


class WinBasedAttention(B200Module):
    """reference win_attention.py:118-207 (always used with shift_size > 0 by WACNN)."""

    def __init__(self, dim=192, num_heads=8, window_size=8, shift_size=0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0):
        super().__init__()
        if not 0 <= shift_size < window_size:
            raise AssertionError("shift_size must in 0-window_size")  # win_attention.py:144
        if drop or attn_drop or drop_path:
            raise ValueError("dropout / drop-path are always 0 in the reference")
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.attn = WindowAttention(dim, window_size=(window_size, window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, qk_scale=qk_scale)

    def emit(self, ctx: Ctx, x, **kw):
        return self.attn.emit(ctx, x, shift=self.shift_size, shortcut=x)
