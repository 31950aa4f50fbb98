"""Golden fixture for the rate-distortion TRAINING STEP (BASELINE config 4: lambda = 0.0035, fp32):
the UNMODIFIED reference WACNN in `.train()` mode, the reference's RateDistortionLoss
(training/loss.py:6-30), `loss.backward()` and `aux_loss().backward()` exactly as training/step.py:42-52 runs them,
on the reference's own random init (seed 0) and a 2 x 3 x 256 x 256 batch.

Stored: the noise draws (re-drawn in the reference's draw order, asserted aligned), the loss terms, and for EVERY
parameter the gradient's L2 norm, its sum and 16 sampled entries (the full gradients are 301 MB).

    python tests/golden/make_golden_rdstep.py      (build container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, wacnn_oracle, weights  # noqa: E402
from tests.golden.make_golden_train import redraw_noise  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, LMBDA = 9876, 0.0035
N_SAMPLES = 16
# case -> (weights, B, H, W): the reference's own init at config 4's image size, and the hash-seeded "stress" weights
# (large activations, symbols up to +-15: every gradient path is exercised) at a smaller size
CASES = {"rdstep_c256": ("refinit", 2, 256, 256), "rdstep_hash_c128": ("hash", 2, 128, 128)}


def sample_index(k, numel):
    return np.random.RandomState(1000 + k).randint(0, numel, size=N_SAMPLES)


def main():
    torch.set_num_threads(8)
    for case, (wkind, B, H, W) in CASES.items():
        run_case(case, wkind, B, H, W)


def run_case(case, wkind, B, H, W):
    sd = weights.refinit_state_dict(0) if wkind == "refinit" else weights.make_state_dict(seed=0)
    net = ref_shim.reference_wacnn()
    net.load_state_dict(sd, strict=True)
    net.train()
    from compress.training.loss import RateDistortionLoss  # the reference's own criterion
    crit = RateDistortionLoss(lmbda=LMBDA)
    x = weights.rand_image(B, H, W, seed=11)
    torch.manual_seed(SEED)
    out = net(x)
    oc = crit(out, x)
    oc["loss"].backward()
    names = [n for n, p in net.named_parameters()]
    grads = {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in net.named_parameters()}
    net.zero_grad()
    aux = net.aux_loss()
    aux.backward()
    aux_grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None and p.grad.abs().sum() > 0}
    assert set(aux_grads) == {"entropy_bottleneck.quantiles"}, sorted(aux_grads)

    torch.manual_seed(SEED)
    noise = redraw_noise(B, H, W)
    with torch.no_grad():
        chk = wacnn_oracle.forward(sd, x, noise=noise)
    for k in ("y", "z"):
        err = (chk["likelihoods"][k] - out["likelihoods"][k]).abs().max().item()
        assert err < 1e-5, (k, err)  # a misaligned noise re-draw would be off by O(0.1)

    norms = np.zeros(len(names), np.float64)
    sums = np.zeros(len(names), np.float64)
    samples = np.zeros((len(names), N_SAMPLES), np.float32)
    has_grad = np.zeros(len(names), np.bool_)
    for k, n in enumerate(names):
        g = grads[n]
        if g is None:
            continue
        has_grad[k] = True
        g64 = g.double().reshape(-1)
        norms[k], sums[k] = g64.norm().item(), g64.sum().item()
        samples[k] = g.reshape(-1)[torch.from_numpy(sample_index(k, g.numel()))].numpy()
    res = dict(names=np.array(names), has_grad=has_grad, grad_norm=norms, grad_sum=sums, grad_samples=samples,
               loss=np.float64(oc["loss"].item()), bpp_loss=np.float64(oc["bpp_loss"].item()),
               mse_loss=np.float64(oc["mse_loss"].item()), aux_loss=np.float64(aux.item()),
               aux_grad_quantiles=aux_grads["entropy_bottleneck.quantiles"].numpy(),
               noise_y=noise["y"].numpy().astype(np.float32), noise_z=noise["z"].numpy().astype(np.float32),
               image_seed=np.array(11), lmbda=np.float64(LMBDA))
    path = os.path.join(HERE, f"wacnn_{case}.npz")
    np.savez_compressed(path, **res)
    print(case, "loss", oc["loss"].item(), "bpp", oc["bpp_loss"].item(), "mse", oc["mse_loss"].item(), "aux", aux.item())
    print("params", len(names), "with grad", int(has_grad.sum()), "total grad norm", float(np.sqrt((norms ** 2).sum())),
          "file", os.path.getsize(path))
    order = np.argsort(norms)
    print("smallest grad norms:", [(names[i], norms[i]) for i in order[:5]])
    print("largest grad norms:", [(names[i], norms[i]) for i in order[-5:]])


if __name__ == "__main__":
    main()
