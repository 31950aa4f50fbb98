"""Parameter containers mirroring the reference's `compress/ops` (bound_ops.py,
parametrizers.py, ops.py).  In the B200 path their arithmetic is fused into the
kernels (GDN packing, GC/EB epilogues); these classes exist so that state_dict
keys (`...beta_reparam.pedestal`, `...lower_bound.bound`) match the reference."""
import torch
import torch.nn as nn


class LowerBound(nn.Module):
    """reference ops/bound_ops.py:44-65: max(x, bound); holds the `bound` buffer."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


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


__all__ = ["LowerBound", "NonNegativeParametrizer"]
