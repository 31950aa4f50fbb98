"""GDN / inverse GDN (reference layers/gdn.py:25-75) as ONE kernel: the 1x1
contraction beta + gamma @ x^2 runs in the implicit-GEMM kernel with the A
operand squared on load, and rsqrt (or sqrt) and the multiply by x are its
epilogue.  The reference launches pow, conv2d, rsqrt and mul separately."""
import torch
import torch
import torch.nn as nn

from .. import _lib, packing
from ..ops import NonNegativeParametrizer
from .base import B200Module, Ctx


class GDN(B200Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.in_channels = in_channels
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def packed(self, wdt):
        return self._packed(("gdn", wdt), (self.beta, self.gamma), lambda: packing.pack_gdn(self.beta, self.gamma, wdt))

    wants_square = True  # in bf16 mode the producing conv also stores x^2 (see Sequential.emit)

    def fused_params(self):
        """(gamma' bf16 [C][C], beta' fp32, inverse) for the fused conv+GDN kernel."""
        gamma, beta = self.packed(torch.bfloat16)
        return gamma, beta, self.inverse

    def emit(self, ctx: Ctx, x, out=None, x2=None, **kw):
        C = self.in_channels
        if out is None:
            out = ctx.buf(x.B, x.H, x.W, C)
        epi = _lib.EPI_IGDN if self.inverse else _lib.EPI_GDN
        if ctx.wdt_for(x) == torch.bfloat16:
            # tensor-core path: TMA cannot square on load, so A = x^2 comes from a side buffer
            if x2 is None:
                x2 = ctx.prog.copy(x, ctx.buf(x.B, x.H, x.W, C), op_code=2)
            gamma, beta = self.packed(torch.bfloat16)
            return ctx.prog.conv(x2, gamma, beta, C, 1, 1, 1, 0, 0, out, epilogue=epi, res=x)
        gamma, beta = self.packed(torch.float32)
        return ctx.prog.conv(x, gamma, beta, C, 1, 1, 1, 0, 0, out, epilogue=epi, res=x, a_square=True)
