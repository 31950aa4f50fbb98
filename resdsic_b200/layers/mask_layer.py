"""ResDSIC importance masks (reference layers/mask_layer.py:6-127): which latent elements of the progressive
stream are transmitted at a given quality level.

Parameters keep the reference's names (`masking.gamma`, `masking.mask_conv.0.weight`, ...).  Evaluation mode is
implemented: `Mask.forward` followed by `apply_noise(mask, tr=False)` = `torch.round` (:32-39), i.e. a 0/1 mask.
The 1x1 `mask_conv` over `cat(scale, scale_prog)` runs in the implicit-GEMM kernel as a split-K pair (the cat is
never materialised); sigmoid / pow / round is one elementwise kernel (`rdsic_mask_forward`).

Policies: "two-levels", "learnable-mask-gamma", "learnable-mask-nested" (as in the reference).  Not provided:
"point-based-std" (a global `torch.quantile` over the whole batch tensor -- a sort, outside the per-image
hot path) and "scalable_res" (raises AttributeError in the reference itself: `self.lmbda_list` is undefined at
mask_layer.py:118).
"""
import torch
import torch.nn as nn

from .. import _lib
from .base import B200Module, Ctx, Sequential
from .conv import Conv2d

POLICIES = ("two-levels", "learnable-mask-gamma", "learnable-mask-nested")

ZEROS, ONES = "zeros", "ones"


class Mask(B200Module):
    def __init__(self, mask_policy, scalable_levels, M):
        super().__init__()
        self.mask_policy, self.scalable_levels, self.M = mask_policy, scalable_levels, M
        if mask_policy == "learnable-mask-gamma":
            self.gamma = nn.Parameter(torch.ones((scalable_levels - 2, M)))
            self.mask_conv = Sequential(Conv2d(2 * M, M, 1, 1))
        if mask_policy == "learnable-mask-nested":
            self.mask_conv = nn.ModuleList(Sequential(Conv2d(2 * M, M, 1, 1)) for _ in range(scalable_levels - 2))

    def kind(self, pr):
        """ZEROS / ONES for the constant cases of Mask.forward (mask_layer.py:64-113), else None (computed)."""
        pol = self.mask_policy
        if pol == "two-levels":
            return ZEROS if pr == 0 else ONES
        if pol == "learnable-mask-gamma":
            return ZEROS if pr == 0 else (ONES if pr == self.scalable_levels - 1 else None)
        if pol == "learnable-mask-nested":
            return ZEROS if pr == 0 else (ONES if pr == 1 else None)
        if pol in ("point-based-std", "scalable_res"):
            raise NotImplementedError(f"mask policy {pol!r} is not provided by resdsic_b200 (see layers/mask_layer.py)")
        raise NotImplementedError(pol)

    def gamma_vector(self, pr, device):
        """relu(sum of the first scalable_levels-1-pr rows of gamma) + 1e-7 (mask_layer.py:78-84), on the host."""
        index_pr = int(self.scalable_levels - 1 - pr)
        g = self.gamma.detach().float().cpu()
        g = torch.relu(torch.sum(torch.stack([g[j] for j in range(index_pr)]), dim=0)) + 1e-7
        return g.contiguous().to(device)

    def emit_logits(self, ctx: Ctx, conv_seq, scale, scale_prog):
        """mask_conv(cat(scale, scale_prog)) -> fp32 logits [B,h,w,M] (split-K over the two halves of the cat)."""
        conv = conv_seq[0]
        part = conv.emit_partial(ctx, scale, 0, self.M)
        return conv.emit_partial(ctx, scale_prog, 1, self.M, res=part, out_dtype=torch.float32)

    def emit(self, ctx: Ctx, scale, scale_prog, pr, cache=None):
        """0/1 mask [B,h,w,M] fp32 for quality index `pr` (must not be a constant case).  `cache` (dict) shares
        the conv logits between quality levels."""
        cache = {} if cache is None else cache
        out = ctx.buf(scale.B, scale.H, scale.W, self.M, torch.float32)
        if self.mask_policy == "learnable-mask-gamma":
            if "g" not in cache:
                cache["g"] = self.emit_logits(ctx, self.mask_conv, scale, scale_prog)
            return ctx.prog.mask([cache["g"]], out, 1, gamma=self.gamma_vector(pr, ctx.device))
        ins = []
        for i in range(pr):
            if i not in cache:
                cache[i] = self.emit_logits(ctx, self.mask_conv[i], scale, scale_prog)
            ins.append(cache[i])
        return ctx.prog.mask(ins, out, 2)
