"""RateDistortionLoss (reference training/loss.py:6-30) on the library's reduction kernels."""
import math

import torch
import torch.nn as nn

from . import functions as Fn


class RateDistortionLoss(nn.Module):
    """loss = lmbda * 255^2 * MSE(x_hat, target) + sum_k sum(log(likelihoods_k)) / (-ln 2 * N * H * W)."""

    def __init__(self, lmbda=0.05):
        super().__init__()
        self.lmbda = lmbda

    def forward(self, output, target, lmbda=None):
        N, _, H, W = target.size()
        num_pixels = N * H * W
        lmbda = self.lmbda if lmbda is None else lmbda
        out = {}
        out["bpp_loss"] = sum(Fn.SumLogFn.apply(lik) / (-math.log(2) * num_pixels) for lik in output["likelihoods"].values())
        out["mse_loss"] = Fn.MseFn.apply(output["x_hat"], target.float())
        out["loss"] = lmbda * 255 ** 2 * out["mse_loss"] + out["bpp_loss"]
        return out
