"""Swin transformer block for the `stf` family.

The reference tree has no STF model; its in-tree statement of the LayerNorm + W/SW-MSA + MLP block
is `Block` / `WMSA` of the unregistered `models/TCM/tcm.py:139-236`, which this module mirrors name for name
(`ln1`, `msa.embedding_layer`, `msa.relative_position_params`, `msa.linear`, `ln2`, `mlp.0`, `mlp.2`), so its
state_dict loads unchanged.  Seven launches: LN, qkv GEMM, fused window attention (roll / partition / mask
are addressing), proj GEMM (+residual), LN, MLP GEMM (+GELU), MLP GEMM (+residual)."""
import torch
import torch.nn as nn

from .. import _lib
from .base import GELU, B200Module, Ctx, Sequential
from .conv import Linear


class LayerNorm(B200Module):
    """nn.LayerNorm over the channel dim of a channels-last view."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def emit(self, ctx: Ctx, x, out=None, out_dtype=None, **kw):
        g, b = self._packed("ln", (self.weight, self.bias), lambda: (self.weight.detach().float().contiguous(),
                                                                    self.bias.detach().float().contiguous()))
        return ctx.prog.layernorm(x, out if out is not None else ctx.buf(x.B, x.H, x.W, x.C, out_dtype), g, b, self.eps)


class WMSA(B200Module):
    """tcm.py:139-212.  `type` 'W' (no shift) or 'SW' (shift = window_size // 2)."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, type):
        super().__init__()
        if input_dim % head_dim:
            raise ValueError("input_dim must be a multiple of head_dim")
        self.input_dim, self.output_dim, self.head_dim = input_dim, output_dim, head_dim
        self.n_heads, self.window_size, self.type = input_dim // head_dim, window_size, type
        self.scale = head_dim ** -0.5
        self.embedding_layer = Linear(input_dim, 3 * input_dim)
        t = torch.zeros((2 * window_size - 1) * (2 * window_size - 1), self.n_heads)
        nn.init.trunc_normal_(t, std=0.02, a=-2.0, b=2.0)
        self.relative_position_params = nn.Parameter(
            t.view(2 * window_size - 1, 2 * window_size - 1, self.n_heads).permute(2, 0, 1).contiguous())
        self.linear = Linear(input_dim, output_dim)

    def bias_table(self):
        p = self.relative_position_params  # [heads, 2ws-1, 2ws-1] -> kernel layout [(2ws-1)^2, heads]
        return self._packed("tab", (p,), lambda: p.detach().float().permute(1, 2, 0).reshape(-1, self.n_heads).contiguous())

    def emit(self, ctx: Ctx, x, shortcut=None, **kw):
        ws = self.window_size
        if x.H % ws or x.W % ws:
            raise ValueError(f"feature map {x.H}x{x.W} is not a multiple of the window size {ws}")
        qkv = self.embedding_layer.emit(ctx, x)
        att = ctx.buf(x.B, x.H, x.W, self.input_dim)
        ctx.prog.attn(qkv, att, self.bias_table(), self.n_heads, ws, ws // 2 if self.type != "W" else 0, float(self.scale))
        if shortcut is not None:
            return self.linear.emit(ctx, att, epilogue=_lib.EPI_ADD_RES, res=shortcut)
        return self.linear.emit(ctx, att)


class Block(B200Module):
    """tcm.py:214-236: x += msa(ln1(x)); x += mlp(ln2(x)).  forward() takes/returns NCHW like every other
    B200 layer; `forward_nhwc` takes the reference's [b,h,w,c] layout."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path=0.0, type="W", input_resolution=None):
        super().__init__()
        if type not in ("W", "SW"):
            raise AssertionError("type must be 'W' or 'SW'")  # tcm.py:221
        if drop_path:
            raise ValueError("drop_path is 0 in every reference configuration")
        if input_dim != output_dim:
            raise ValueError("the residual form needs output_dim == input_dim")
        self.input_dim, self.output_dim, self.type = input_dim, output_dim, type
        self.ln1 = LayerNorm(input_dim)
        self.msa = WMSA(input_dim, input_dim, head_dim, window_size, type)
        self.ln2 = LayerNorm(input_dim)
        self.mlp = Sequential(Linear(input_dim, 4 * input_dim), GELU(), Linear(4 * input_dim, output_dim))

    def emit(self, ctx: Ctx, x, **kw):
        """`kw` (out / out_dtype / out2) places the block's output (used by the last block of a transform)."""
        x = self.msa.emit(ctx, self.ln1.emit(ctx, x), shortcut=x)
        h = self.mlp[0].emit(ctx, self.ln2.emit(ctx, x), epilogue=_lib.EPI_GELU)
        return self.mlp[2].emit(ctx, h, epilogue=_lib.EPI_ADD_RES, res=x, **kw)

    @torch.no_grad()
    def forward_nhwc(self, x):
        return self.forward(x.permute(0, 3, 1, 2).contiguous()).permute(0, 2, 3, 1).contiguous()


class PatchMerging(B200Module):
    """Swin patch merging (2x down): gather each 2x2 neighbourhood into 4C channels, LayerNorm(4C),
    Linear(4C -> 2C, no bias).  The gather is the im2col kernel (k = 2, stride 2), i.e. channel order
    (dy, dx) = (0,0), (0,1), (1,0), (1,1) -- builder-defined (the reference tree has no STF)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.norm = LayerNorm(4 * dim)
        self.reduction = Linear(4 * dim, 2 * dim, bias=False)

    def emit(self, ctx: Ctx, x, **kw):
        if x.H % 2 or x.W % 2:
            raise ValueError("PatchMerging needs even spatial dims")
        g = ctx.prog.patchify(x, ctx.buf(x.B, x.H // 2, x.W // 2, 4 * self.dim), 2, 2, 2, 0)
        return self.reduction.emit(ctx, self.norm.emit(ctx, g), **kw)


class PatchSplit(B200Module):
    """Inverse of PatchMerging for the synthesis transform (2x up): Linear(C -> 2C, no bias), PixelShuffle(2)
    folded into the GEMM's store addressing (-> C/2 channels), LayerNorm(C/2).  Builder-defined."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.reduction = Linear(dim, 2 * dim, bias=False)
        self.norm = LayerNorm(dim // 2)

    def emit(self, ctx: Ctx, x, **kw):
        up = self.reduction.emit(ctx, x, out=ctx.buf(x.B, 2 * x.H, 2 * x.W, self.dim // 2), pixel_shuffle=2)
        return self.norm.emit(ctx, up, **kw)


class BasicLayer(B200Module):
    """A stage of `depth` Swin blocks alternating W-MSA / SW-MSA, optionally followed by a resampler."""

    def __init__(self, dim, depth, num_heads, window_size, resample=None):
        super().__init__()
        self.blocks = nn.ModuleList(
            Block(dim, dim, dim // num_heads, window_size, 0.0, "W" if i % 2 == 0 else "SW") for i in range(depth))
        if resample == "down":
            self.downsample = PatchMerging(dim)
        elif resample == "up":
            self.upsample = PatchSplit(dim)
        self.resample = resample

    def emit(self, ctx: Ctx, x, **kw):
        blocks = list(self.blocks)
        for i, blk in enumerate(blocks):
            last = i + 1 == len(blocks) and self.resample is None
            x = blk.emit(ctx, x, **(kw if last else {}))
        if self.resample == "down":
            x = self.downsample.emit(ctx, x, **kw)
        elif self.resample == "up":
            x = self.upsample.emit(ctx, x, **kw)
        return x
