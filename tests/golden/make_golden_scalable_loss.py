"""Golden for `ScalableRateDistortionLoss` (reference training/loss.py:33-89): the UNMODIFIED reference criterion
evaluated on the reference's own scalable forward outputs (tests/golden/scalable_icd_gamma.npz) and the image that
produced them, plus the gradients its `loss.backward()` sends into `x_hat` and the four likelihood tensors.

    python tests/golden/make_golden_scalable_loss.py        (build container only)

Output: tests/golden/scalable_loss.npz (scalars + per-tensor gradient checksums).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASE, IMAGE, LMBDA = "icd_gamma", (1, 64, 128), [0.0035, 0.02, 0.065]
KEYS = ("x_hat", "lik_y", "lik_z", "lik_z_prog", "lik_y_prog")


def load_case():
    g = np.load(os.path.join(HERE, f"scalable_{CASE}.npz"))
    t = {k: torch.from_numpy(g[k]).clone() for k in KEYS}
    return t, weights.make_image(*IMAGE, seed=5)


def main():
    ref_shim.install()
    from compress.training.loss import ScalableRateDistortionLoss
    t, x = load_case()
    for v in t.values():
        v.requires_grad_(True)
    out = {"x_hat": t["x_hat"], "likelihoods": {"y": t["lik_y"], "z": t["lik_z"], "z_prog": t["lik_z_prog"], "y_prog": t["lik_y_prog"]}}
    crit = ScalableRateDistortionLoss(lmbda_list=LMBDA, device="cpu")
    res = crit(out, x)
    res["loss"].backward()
    save = {k: v.detach().numpy() for k, v in res.items()}
    for k, v in t.items():
        save[f"grad_sum_{k}"] = v.grad.double().sum().numpy()
        save[f"grad_l2_{k}"] = v.grad.double().norm().numpy()
    np.savez(os.path.join(HERE, "scalable_loss.npz"), **save)
    print({k: (v.tolist() if v.size < 4 else v.shape) for k, v in save.items()})


if __name__ == "__main__":
    main()
