"""Golden fixtures for the ResDSIC scalable models: run the UNMODIFIED reference
`scalable_icd` / `scalable_imd` / `conditional_scalable_icd` / `conditional_scalable_imd` /
`ResWACNNIndependentEntropy` (imported through oracle/ref_shim.py) in eval mode.

    python tests/golden/make_golden_scalable.py        (build container only)

Weights: hash-seeded values for every entry of the reference model's own state_dict
(`oracle.weights.synth_state_dict`, non-degenerate latents), images from `make_image`.
Also asserts that this repo's constructors draw the reference constructors' init bit for bit under one seed.
Outputs: tests/golden/scalable_<case>.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# name: (registry key, constructor kwargs, quality list passed to forward (None = every lambda), image B,H,W)
CASES = {
    "icd_gamma": ("icd", dict(lambda_list=[0.0035, 0.02, 0.065], mask_policy="learnable-mask-gamma", lrp_prog=True), None, (1, 64, 128)),
    "imd_two": ("imd", dict(lambda_list=[0.0035, 0.065], mask_policy="two-levels", lrp_prog=True), None, (2, 64, 64)),
    "icd_nested": ("icd", dict(lambda_list=[0.0035, 0.01, 0.02, 0.065], mask_policy="learnable-mask-nested", lrp_prog=True,
                               independent_lrp=True), [0.01, 0.02], (1, 64, 64)),  # (pr = levels-1 indexes past mask_conv in the reference)
    "icd_nolrp": ("icd", dict(lambda_list=[0.0035, 0.065], mask_policy="two-levels", lrp_prog=False), [0.065], (1, 64, 64)),
    "cicd_two": ("cicd", dict(lambda_list=[0.0035, 0.065], mask_policy="two-levels", lrp_prog=True, joiner_policy="conditional"),
                 None, (1, 64, 64)),
    # cimd: computed (learnable) masks ARE reachable, likelihood at scale * mask + 1e-7, two decoders, no "y" output
    "cimd_gamma": ("cimd", dict(lambda_list=[0.0035, 0.02, 0.065], mask_policy="learnable-mask-gamma", lrp_prog=True,
                                joiner_policy="conditional"), None, (1, 64, 64)),
    "cimd_res": ("cimd", dict(lambda_list=[0.0035, 0.065], mask_policy="two-levels", lrp_prog=True, joiner_policy="residual"),
                 [0.065], (1, 64, 64)),
    "cimd_cat": ("cimd", dict(lambda_list=[0.0035, 0.065], mask_policy="two-levels", lrp_prog=True, joiner_policy="concatenation"),
                 None, (1, 64, 64)),  # 2M-wide enhancement decoder fed with cat(base, progressive)
    # ind: independent entropy models, no mask on the data path; single decoder / decoder pair + independent LRP
    "ind_two": ("ind", dict(lambda_list=[0.0035, 0.065], mask_policy="two-levels", lrp_prog=True), None, (1, 64, 64)),
    "ind_md": ("ind", dict(lambda_list=[0.0035, 0.065], mask_policy="learnable-mask", lrp_prog=True, independent_lrp=True,
                           multiple_decoder=True), None, (1, 64, 64)),
}
GAINS = ((r"^g_(a|s|a_progressive)\.", 0.7071), (r"^g_a(_progressive)?\.7\.weight$", 2.5), (r"^h_a(_prog)?\.8\.weight$", 3.0),
                              (r"^h_(mean|scale)_s(_prog)?\.8\.weight$", 1.5), (r"^cc_scale_transforms(_prog)?\.\d\.8\.weight$", 1.5),
                              (r"^cc_mean_transforms(_prog)?\.\d\.8\.weight$", 0.4), (r"^g_s(\.\d)?\.[136]\.weight$", 0.25),
                              (r"^g_s(\.\d)?\.8\.weight$", 0.35), (r"^masking\.mask_conv\.", 3.0))


def case_state_dict(ref_sd):
    """Hash-seeded weights keyed on the model's own state_dict; GDN parameters keep the reference's init values."""
    sd = weights.synth_state_dict({k: v for k, v in ref_sd.items() if not k.endswith((".beta", ".gamma")) and k != "gamma"},
                                  seed=0, gains=GAINS)
    out = {}
    for k, v in ref_sd.items():
        if k == "masking.gamma":
            out[k] = 1.0 + weights.hash_symmetric(k, tuple(v.shape), 0.8)
        elif k in sd and not k.endswith(("_prog.scale_table",)):
            out[k] = sd[k]
        else:
            out[k] = v.clone()
    return out


def main(only=None):
    import resdsic_b200
    torch.set_num_threads(8)
    ref_shim.install()
    from compress.models import models as RM
    for name, (key, kw, quality, (B, H, W)) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        net = RM[key](N=192, M=320, **kw).eval()
        torch.manual_seed(0)
        mine = resdsic_b200.models[key](N=192, M=320, **kw)
        a, b = net.state_dict(), mine.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), f"{name}: constructor init differs at {k}"
        sd = case_state_dict(a)
        torch.nn.Module.load_state_dict(net, sd, strict=True)  # (the model's own load_state_dict resizes CDF buffers first)
        x = weights.make_image(B, H, W, seed=5)
        with torch.no_grad():
            out = net(x, quality=quality)
        res = {"x_hat": out["x_hat"], "lik_y": out["likelihoods"]["y"], "lik_z": out["likelihoods"]["z"],
               "lik_z_prog": out["likelihoods"]["z_prog"], "lik_y_prog": out["likelihoods"]["y_prog"],
               "z_hat": out["z_hat"], "z_hat_prog": out["z_hat_prog"]}
        if "y" in out:  # (cimd's forward does not return it)
            res["y"] = out["y"]
        res = {k: v.detach().numpy() for k, v in res.items()}
        qs = [net.lmbda_index_list[q] for q in (quality if quality is not None else net.lmbda_list)]
        res["qualities"] = np.array(qs)
        path = os.path.join(HERE, f"scalable_{name}.npz")
        np.savez_compressed(path, **res)
        print(name, {k: v.shape for k, v in res.items()}, "x_hat range", res["x_hat"].min(), res["x_hat"].max(),
              "lik_y_prog min", res["lik_y_prog"].min(), os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv[1:])
