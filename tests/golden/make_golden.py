"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (`/root/reference`, imported through oracle/ref_shim.py) on the
deterministic synthetic weights/images of oracle/weights.py.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/wacnn_<case>.npz, tests/golden/ops.npz.
The arithmetic executed here is the reference's own modules
(`compress.models.WACNN.cnn.WACNN` and its sub-modules); nothing from the
oracle's restatement or from resdsic_b200 is on the path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"c64": (1, 64, 64), "c128x192": (2, 128, 192)}
LOWRATE_CASES = {"lowrate_c128x192": (2, 128, 192)}  # same shapes, weights of profile "lowrate"


def checksum(sd):
    """A few exact numbers that pin the generated weights."""
    names = ["g_a.0.weight", "g_a.4.conv_b.0.attn.qkv.weight", "g_s.6.weight", "lrp_transforms.9.0.weight",
             "entropy_bottleneck._matrix2", "g_a.3.gamma"]
    out = {}
    for n in names:
        t = sd[n].double().reshape(-1)
        out[n] = np.array([t.sum().item(), t[:: max(1, t.numel() // 7)][:7].abs().sum().item(), float(t[-1])])
    return out


@torch.no_grad()
def run_reference_model(net, x, table):
    """Reference forward + the symbol/index build of `compress` (cnn.py:217-268),
    executed with the reference's own sub-modules."""
    grabbed = {}

    def grab(name):
        def hook(_m, _i, o):
            grabbed[name] = o.detach().clone()
        return hook

    hooks = [net.g_a.register_forward_hook(grab("y")), net.h_a.register_forward_hook(grab("z")),
             net.h_mean_s.register_forward_hook(grab("latent_means")),
             net.h_scale_s.register_forward_hook(grab("latent_scales"))]
    out = net(x)
    for h in hooks:
        h.remove()
    y, means, scales = grabbed["y"], grabbed["latent_means"], grabbed["latent_scales"]
    # z_hat exactly as cnn.py:152-154
    off = net.entropy_bottleneck._get_medians()
    from compress.ops import ste_round
    z_hat = ste_round(grabbed["z"] - off) + off
    # slice loop as in compress(), cnn.py:241-268, with the reference modules
    gc = net.gaussian_conditional
    gc.scale_table = table
    y_hat_slices, syms, idxs, mus, scs = [], [], [], [], []
    for i, y_i in enumerate(y.chunk(net.num_slices, 1)):
        support = y_hat_slices[: net.max_support_slices]
        mean_support = torch.cat([means] + support, 1)
        mu = net.cc_mean_transforms[i](mean_support)
        scale = net.cc_scale_transforms[i](torch.cat([scales] + support, 1))
        idxs.append(gc.build_indexes(scale))
        q = gc.quantize(y_i, "symbols", mu)
        syms.append(q)
        y_hat_i = q + mu
        lrp = net.lrp_transforms[i](torch.cat([mean_support, y_hat_i], 1))
        y_hat_i = y_hat_i + 0.5 * torch.tanh(lrp)
        y_hat_slices.append(y_hat_i)
        mus.append(mu)
        scs.append(scale)
    res = dict(x_hat=out["x_hat"], lik_y=out["likelihoods"]["y"], lik_z=out["likelihoods"]["z"],
               y=y, z=grabbed["z"], z_hat=z_hat, latent_means=means, latent_scales=scales,
               y_hat=torch.cat(y_hat_slices, 1), mu=torch.cat(mus, 1), scale=torch.cat(scs, 1),
               symbols=torch.cat(syms, 1), indexes=torch.cat(idxs, 1))
    return {k: v.numpy() for k, v in res.items()}


def op_inputs():
    """Inputs of the op-level fixtures; tests regenerate them with the same calls."""
    hs = weights.hash_symmetric
    gc_y = hs("ops.gc.y", (2, 32, 8, 12), 12.0)
    gc_mu = hs("ops.gc.mu", (2, 32, 8, 12), 3.0)
    gc_scale = weights.hash_uniform("ops.gc.scale", (2, 32, 8, 12)) ** 4 * 300.0 - 0.05
    # edge cases: exact .5 ties (round-half-even), scale exactly on table entries / below the bound
    gc_y.view(-1)[:8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5, -3.5])
    gc_mu.view(-1)[:8] = 0.0
    tab = weights.scale_table()
    gc_scale.view(-1)[8:72] = tab
    gc_scale.view(-1)[72:76] = torch.tensor([0.0, -1.0, 0.11, 1e-3])
    return dict(
        gc_y=gc_y, gc_mu=gc_mu, gc_scale=gc_scale,
        eb_z=hs("ops.eb.z", (2, 192, 3, 5), 6.0),
        attn8_x=hs("ops.attn8.x", (1, 192, 16, 24), 1.5),
        attn4_x=hs("ops.attn4.x", (2, 320, 8, 4), 1.5),
        gdn_x=hs("ops.gdn.x", (1, 192, 6, 5), 2.0),
        deconv_x=hs("ops.deconv.x", (1, 192, 5, 4), 1.0),
        block8_x=hs("ops.block8.x", (1, 192, 8, 16), 1.0),
    )


@torch.no_grad()
def run_reference_ops(net, table):
    i = op_inputs()
    gc = net.gaussian_conditional
    gc.scale_table = table
    o = {}
    _, o["gc_lik"] = gc(i["gc_y"], i["gc_scale"], i["gc_mu"])
    o["gc_yhat"] = gc.quantize(i["gc_y"], "dequantize", i["gc_mu"])
    o["gc_sym"] = gc.quantize(i["gc_y"], "symbols", i["gc_mu"])
    o["gc_idx"] = gc.build_indexes(i["gc_scale"])
    o["eb_zhat"], o["eb_lik"] = net.entropy_bottleneck(i["eb_z"])
    o["attn8"] = net.g_a[4].conv_b[0](i["attn8_x"])
    o["attn4"] = net.g_a[8].conv_b[0](i["attn4_x"])
    o["gdn"] = net.g_a[1](i["gdn_x"])
    o["igdn"] = net.g_s[2](i["gdn_x"])
    o["deconv"] = net.g_s[3](i["deconv_x"])
    o["block8"] = net.g_a[4](i["block8_x"])
    o["ru"] = net.g_a[4].conv_a[0](i["block8_x"])
    return {k: v.numpy() for k, v in o.items()}


SWIN_CASES = {  # name: (dim, head_dim, window, type, B, H, W)
    "swin_w8_sw": (192, 24, 8, "SW", 1, 16, 24),
    "swin_w4_w": (128, 32, 4, "W", 2, 8, 12),
    "swin_w4_sw": (128, 32, 4, "SW", 2, 8, 12),
}


def swin_state_dict(name, dim, head_dim, ws):
    """Hash-seeded parameters with the key names of the reference's tcm.Block."""
    hs = weights.hash_symmetric
    heads = dim // head_dim
    sd = {}
    for ln in ("ln1", "ln2"):
        sd[f"{ln}.weight"] = 1.0 + hs(f"{name}.{ln}.w", (dim,), 0.3)
        sd[f"{ln}.bias"] = hs(f"{name}.{ln}.b", (dim,), 0.2)
    sd["msa.relative_position_params"] = hs(f"{name}.rpp", (heads, 2 * ws - 1, 2 * ws - 1), 0.5)
    for lin, (o, i) in (("msa.embedding_layer", (3 * dim, dim)), ("msa.linear", (dim, dim)),
                        ("mlp.0", (4 * dim, dim)), ("mlp.2", (dim, 4 * dim))):
        sd[f"{lin}.weight"] = hs(f"{name}.{lin}.w", (o, i), (3.0 / i) ** 0.5)
        sd[f"{lin}.bias"] = hs(f"{name}.{lin}.b", (o,), 0.1)
    return sd


@torch.no_grad()
def run_reference_swin():
    """Outputs of the reference's own `Block` (models/TCM/tcm.py:214-236), NHWC in / NHWC out."""
    tcm = ref_shim.reference_tcm_module()
    out = {}
    for name, (dim, hd, ws, typ, B, H, W) in SWIN_CASES.items():
        blk = tcm.Block(dim, dim, hd, ws, 0.0, typ).eval()
        sd = swin_state_dict(name, dim, hd, ws)
        assert set(sd) == set(blk.state_dict()), (sorted(sd), sorted(blk.state_dict()))
        blk.load_state_dict(sd, strict=True)
        x = weights.hash_symmetric(f"{name}.x", (B, H, W, dim), 1.5)
        out[name] = blk(x).numpy()
    return out


def main():
    torch.set_num_threads(8)
    sd = weights.make_state_dict(seed=0)
    net = ref_shim.reference_wacnn().eval()
    ref_sd = net.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "state_dict inventory differs from the reference"
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape) and ref_sd[k].dtype == sd[k].dtype, k
        if k.endswith(("pedestal", "bound", "target", "relative_position_index")):
            assert torch.equal(ref_sd[k], sd[k]), f"constant buffer {k} differs from the reference's"
    net.load_state_dict(sd, strict=True)
    table = weights.scale_table()
    from compress.models.WACNN.cnn import get_scale_table
    assert torch.equal(table, get_scale_table())
    ck = checksum(sd)
    for case, (B, H, W) in CASES.items():
        x = weights.make_image(B, H, W, seed=0)
        res = run_reference_model(net, x, table)
        res["scale_table"] = table.numpy()
        for n, v in ck.items():
            res["ck:" + n] = v
        path = os.path.join(HERE, f"wacnn_{case}.npz")
        np.savez_compressed(path, **res)
        bpp = sum(np.log(res[k]).sum() for k in ("lik_y", "lik_z")) / (-np.log(2) * B * H * W)
        print(case, "bpp", bpp, "x_hat range", res["x_hat"].min(), res["x_hat"].max(),
              "sym range", res["symbols"].min(), res["symbols"].max(),
              "idx range", res["indexes"].min(), res["indexes"].max(), os.path.getsize(path))
    ops = run_reference_ops(net, table)
    ops.update(run_reference_swin())
    # second operating point: "lowrate" weights (most symbols zero, like a trained codec)
    net.gaussian_conditional.scale_table = torch.Tensor()  # back to the constructor state (was assigned above)
    net.load_state_dict(weights.make_state_dict(seed=0, profile="lowrate"), strict=True)
    for case, (B, H, W) in LOWRATE_CASES.items():
        x = weights.make_image(B, H, W, seed=0)
        res = run_reference_model(net, x, table)
        keep = {k: res[k] for k in ("x_hat", "lik_y", "lik_z", "y", "z", "symbols", "indexes", "y_hat", "latent_means",
                                    "latent_scales")}
        np.savez_compressed(os.path.join(HERE, f"wacnn_{case}.npz"), **keep)
        bpp = sum(np.log(res[k]).sum() for k in ("lik_y", "lik_z")) / (-np.log(2) * B * H * W)
        print(case, "bpp", bpp, "zero symbols", (res["symbols"] == 0).mean())
    path = os.path.join(HERE, "ops.npz")
    np.savez_compressed(path, **ops)
    print("ops", {k: v.shape for k, v in ops.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
