"""RateDistortionLoss / ScalableRateDistortionLoss (reference training/loss.py:6-30,33-89) on the library's
reduction kernels (fp64 accumulation, fixed summation order)."""
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


class ScalableRateDistortionLoss(nn.Module):
    """reference training/loss.py:33-89: the criterion of the ResDSIC scalable models (`x_hat` [levels,B,3,H,W]).

        mse_l   = mean((target - x_hat[l])^2)                                  per quality level
        bpp_*   = sum(log(lik_*)) / (-ln 2 * B*H*W)     for z, y (base) and z_prog, y_prog (scalable)
        bpp_loss = bpp_scalable + levels * bpp_base
        loss    = bpp_loss + weight * mean_l(lmbda[l] * mse_l)
    """

    def __init__(self, weight=255 ** 2, lmbda_list=(0.75,), device="cuda"):
        super().__init__()
        self.scalable_levels = len(lmbda_list)
        self.lmbda = torch.tensor(list(lmbda_list)).to(device)
        self.weight = weight

    def forward(self, output, target, lmbda=None):
        B, _, H, W = target.size()
        den = -math.log(2) * B * H * W
        x_hat = output["x_hat"]
        levels = x_hat.shape[0]
        lmbda = self.lmbda if lmbda is None else torch.tensor(lmbda).to(self.lmbda.device)
        tgt = target.float()
        out = {}
        out["mse_loss"] = torch.stack([Fn.MseFn.apply(x_hat[l], tgt) for l in range(levels)])
        lik = output["likelihoods"]
        out["bpp_hype_base"] = Fn.SumLogFn.apply(lik["z"]) / den
        out["bpp_main_base"] = Fn.SumLogFn.apply(lik["y"]) / den
        out["bpp_base"] = out["bpp_main_base"] + out["bpp_hype_base"]
        out["bpp_hype_scale"] = Fn.SumLogFn.apply(lik["z_prog"]) / den
        out["bpp_main_scale"] = Fn.SumLogFn.apply(lik["y_prog"]) / den
        out["bpp_scalable"] = out["bpp_main_scale"] + out["bpp_hype_scale"]
        out["bpp_loss"] = out["bpp_scalable"] + levels * out["bpp_base"]
        out["loss"] = out["bpp_loss"] + self.weight * (lmbda * out["mse_loss"]).mean()
        return out
