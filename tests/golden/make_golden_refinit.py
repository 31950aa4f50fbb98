"""Golden fixtures on the reference's OWN random init (BASELINE.json north_star: "identical inputs and
random-init weights"; BASELINE configs 1 and 3).

Run in the build container only:

    python tests/golden/make_golden_refinit.py

1. `torch.manual_seed(0); compress.models.WACNN()` -- the unmodified reference constructor -- and
   `resdsic_b200.utils.synthetic.refinit_model(0)` must agree on all 585 state_dict entries bit for bit
   (asserted here); per-tensor checksums of a spread of entries are committed so that the GPU box, which has no
   reference, can verify it re-created the same weights.
2. The reference forward (+ the symbol / index build of `compress`) on `torch.rand` images:
   config 1 = 2x3x256x256 (full tensors committed) and config 3's shape 1x3x512x768 (x_hat sub-sampled 4x4,
   likelihoods in full, per-image bpp / PSNR).
Outputs: tests/golden/refinit_c256.npz, tests/golden/refinit_kodak.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import ref_shim, weights  # noqa: E402
import make_golden as MG  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CK_NAMES = ["g_a.0.weight", "g_a.0.bias", "g_a.4.conv_b.0.attn.qkv.weight", "g_a.4.conv_b.0.attn.relative_position_bias_table",
            "g_a.8.conv_a.2.conv.2.weight", "g_s.1.weight", "g_s.6.bias", "g_s.8.weight", "h_a.8.weight",
            "h_mean_s.2.0.weight", "h_scale_s.8.bias", "cc_mean_transforms.0.0.weight", "cc_scale_transforms.9.8.weight",
            "lrp_transforms.9.0.weight", "lrp_transforms.4.8.bias", "entropy_bottleneck._matrix2",
            "entropy_bottleneck.quantiles", "g_a.3.gamma", "g_s.2.beta"]


def stats(res, x):
    B = x.shape[0]
    n = x.shape[2] * x.shape[3]
    bpp = np.array([sum(np.log(res[k][b].astype(np.float64)).sum() for k in ("lik_y", "lik_z")) / (-np.log(2) * n)
                    for b in range(B)])
    psnr = np.array([-10 * np.log10(((res["x_hat"][b].astype(np.float64) - x[b].numpy()) ** 2).mean()) for b in range(B)])
    return bpp, psnr


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    net = ref_shim.reference_wacnn().eval()  # the reference's constructor under seed 0
    ref_sd = net.state_dict()
    mine = weights.refinit_state_dict(0)
    assert list(ref_sd) == list(mine)
    for k in ref_sd:
        assert ref_sd[k].dtype == mine[k].dtype and torch.equal(ref_sd[k], mine[k]), f"constructor init differs: {k}"
    print("constructor init: all", len(ref_sd), "entries bit-equal to the reference's")
    ck = {"ck:" + n: weights.tensor_checksum(ref_sd[n]) for n in CK_NAMES}
    total = sum(float(v.double().sum()) for v in ref_sd.values() if v.is_floating_point())
    ck["ck:__total__"] = np.array([total])
    table = weights.scale_table()
    for name, (B, H, W) in {"c256": (2, 256, 256), "kodak": (1, 512, 768)}.items():
        x = weights.rand_image(B, H, W, seed=1)
        res = MG.run_reference_model(net, x, table)
        net.gaussian_conditional.scale_table = torch.Tensor()
        bpp, psnr = stats(res, x)
        res["bpp"], res["psnr"] = bpp, psnr
        res["x_ck"] = weights.tensor_checksum(x)
        for k in ("latent_means", "latent_scales", "y_hat", "z_hat"):  # keep the fixtures small
            res.pop(k)
        if name == "kodak":  # x_hat sub-sampled 4x4, float tensors of the slice loop dropped
            res["x_hat_sub"] = res.pop("x_hat")[:, :, ::4, ::4].copy()
            for k in ("mu", "scale"):
                res.pop(k)
        res.update(ck)
        path = os.path.join(HERE, f"refinit_{name}.npz")
        np.savez_compressed(path, **res)
        print(name, "bpp", bpp, "psnr", psnr, "zero symbols", (res["symbols"] == 0).mean(), "max|y|", np.abs(res["y"]).max(),
              os.path.getsize(path))


if __name__ == "__main__":
    main()
