"""Golden fixture for the TRAINING-mode forward values (BASELINE config 4, forward half):
the UNMODIFIED reference WACNN in `.train()` mode (noise quantisation inside
EntropyBottleneck / GaussianConditional, entropy_models.py:131-137,447-490,646-661) and
`aux_loss()` (WACNN/base.py:22-27), on the hash-seeded weights of oracle/weights.py.

The reference draws its noise internally with `uniform_`; this script re-draws the same
tensors from the same torch seed in the reference's draw order (one [C,1,B*h*w] tensor for
the bottleneck, then one [B,32,h,w] tensor per slice) and stores them, so that the oracle
and the CUDA path can be fed identical noise.  That the re-draw is aligned is asserted
below (the oracle fed with it reproduces the reference's likelihoods).

    python tests/golden/make_golden_train.py      (build container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, wacnn_oracle, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 4321
CASES = {"train_c64x128": (2, 64, 128)}


def redraw_noise(B, H, W):
    h, w = H // 16, W // 16
    hz, wz = h // 4, w // 4
    nz = torch.empty(192, 1, B * hz * wz).uniform_(-0.5, 0.5)
    ny = [torch.empty(B, 32, h, w).uniform_(-0.5, 0.5) for _ in range(10)]
    return {"z": nz.reshape(192, B, hz, wz).permute(1, 0, 2, 3).contiguous(), "y": torch.cat(ny, 1)}


def main():
    torch.set_num_threads(8)
    sd = weights.make_state_dict(seed=0)
    net = ref_shim.reference_wacnn()
    net.load_state_dict(sd, strict=True)
    net.train()
    for case, (B, H, W) in CASES.items():
        x = weights.make_image(B, H, W, seed=3)
        torch.manual_seed(SEED)
        with torch.no_grad():
            out = net(x)
            aux = net.aux_loss()
        torch.manual_seed(SEED)
        noise = redraw_noise(B, H, W)
        chk = wacnn_oracle.forward(sd, x, noise=noise)
        for k in ("y", "z"):
            err = (chk["likelihoods"][k] - out["likelihoods"][k]).abs().max().item()
            assert err < 1e-5, (k, err)  # a misaligned draw would be off by O(0.1)
        assert (chk["x_hat"] - out["x_hat"]).abs().max().item() < 1e-4
        assert abs(wacnn_oracle.eb_aux_loss(sd).item() - aux.item()) <= 1e-4 * abs(aux.item())
        res = dict(noise_y=noise["y"].numpy(), noise_z=noise["z"].numpy(), x_hat=out["x_hat"].numpy(),
                   lik_y=out["likelihoods"]["y"].numpy(), lik_z=out["likelihoods"]["z"].numpy(),
                   aux_loss=np.array(aux.item(), dtype=np.float64), image_seed=np.array(3))
        path = os.path.join(HERE, f"wacnn_{case}.npz")
        np.savez_compressed(path, **res)
        n = B * H * W
        bpp = sum(np.log(res[k]).sum() for k in ("lik_y", "lik_z")) / (-np.log(2) * n)
        print(case, "train-mode bpp", bpp, "aux_loss", aux.item(), os.path.getsize(path))


if __name__ == "__main__":
    main()
