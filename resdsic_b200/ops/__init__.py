"""Mirror of the reference's `compress/ops` (bound_ops.py, parametrizers.py, ops.py).

Inference: the arithmetic of these operators is fused into the kernels (GDN packing, GC / EB epilogues) and the
classes only carry the reference's state_dict keys (`...beta_reparam.pedestal`, `...lower_bound.bound`).
Training (resdsic_b200/training): `LowerBound` / `NonNegativeParametrizer` are applied to PARAMETER-sized tensors
(GDN beta / gamma) with the reference's custom gradient rule; the activation-sized uses of the same rule (scale
bound 0.11, likelihood bound 1e-9) live in the backward kernels (csrc/train_bwd.cu: gc_backward / eb_backward).
"""
import torch
import torch.nn as nn


class LowerBoundFunction(torch.autograd.Function):
    """reference ops/bound_ops.py:21-42: max(x, bound); the gradient passes where x >= bound or where it is negative
    (i.e. where the update would move x towards the bound)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (grad_output < 0)) * grad_output, None


class LowerBound(nn.Module):
    """reference ops/bound_ops.py:44-65."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return LowerBoundFunction.apply(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """reference ops/parametrizers.py:23-49."""

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


def ste_round(x):
    """reference ops/ops.py:20-34: round with an identity gradient (forward value == torch.round(x) in fp32)."""
    return torch.round(x) - x.detach() + x


__all__ = ["LowerBound", "LowerBoundFunction", "NonNegativeParametrizer", "ste_round"]
