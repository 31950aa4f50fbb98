"""Pins the CPU oracle (oracle/wacnn_oracle.py) against golden vectors produced
by the UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import wacnn_oracle as O
from oracle import weights
from tests.conftest import GOLDEN
from tests.golden.make_golden import CASES, checksum, op_inputs


def _npz(name):
    return np.load(os.path.join(GOLDEN, name))


def test_state_dict_inventory():
    spec = weights.state_dict_spec()
    assert len(spec) == 585
    n_params = sum(int(np.prod(s)) for k, s in spec.items()
                   if not k.endswith(("pedestal", "bound", "target", "relative_position_index"))
                   and not k.rsplit(".", 1)[-1] in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table", "scale_bound"))
    assert n_params == 75_235_779


def test_weight_generator_is_pinned(synthetic_sd):
    g = _npz("wacnn_c64.npz")
    for name, v in checksum(synthetic_sd).items():
        np.testing.assert_array_equal(g["ck:" + name], v, err_msg=name)


def test_scale_table_bit_exact(scale_table):
    np.testing.assert_array_equal(_npz("wacnn_c64.npz")["scale_table"], scale_table.numpy())


@pytest.mark.parametrize("case", list(CASES))
def test_forward_matches_reference(case, synthetic_sd, scale_table):
    B, H, W = CASES[case]
    g = _npz(f"wacnn_{case}.npz")
    x = weights.make_image(B, H, W, seed=0)
    out = O.forward(synthetic_sd, x, scale_table, collect=True)
    # same machine + same ATen conv kernels => expect (near) bit equality;
    # tolerances cover a different host CPU picking other mkldnn kernels.
    for key, ref in (("y", g["y"]), ("z", g["z"]), ("latent_means", g["latent_means"]),
                     ("latent_scales", g["latent_scales"]), ("mu", g["mu"]), ("scale", g["scale"]),
                     ("y_hat", g["y_hat"]), ("x_hat", g["x_hat"])):
        np.testing.assert_allclose(out[key].numpy(), ref, rtol=1e-4, atol=1e-4, err_msg=key)
    np.testing.assert_allclose(out["likelihoods"]["z"].numpy(), g["lik_z"], rtol=1e-4, atol=1e-9)
    # integer outputs: bit-exact except where an upstream 1e-6 wobble crosses a tie
    for key in ("symbols", "indexes"):
        mism = (out[key].numpy() != g[key]).mean()
        assert mism <= 1e-3, (key, mism)
    lik_bad = np.abs(out["likelihoods"]["y"].numpy() - g["lik_y"]) > 1e-4 + 1e-3 * g["lik_y"]
    assert lik_bad.mean() <= 1e-3
    n = B * H * W
    bpp_ref = sum(np.log(g[k].astype(np.float64)).sum() for k in ("lik_y", "lik_z")) / (-np.log(2) * n)
    assert abs(float(O.bpp(out, n)) - bpp_ref) <= 1e-3 * bpp_ref


def test_ops_match_reference(synthetic_sd, scale_table):
    sd, g, i = synthetic_sd, _npz("ops.npz"), op_inputs()
    y_hat = torch.round(i["gc_y"] - i["gc_mu"]) + i["gc_mu"]
    np.testing.assert_array_equal(y_hat.numpy(), g["gc_yhat"])
    np.testing.assert_array_equal(O.gc_symbols(i["gc_y"], i["gc_mu"]).numpy(), g["gc_sym"])
    np.testing.assert_array_equal(O.gc_indexes(i["gc_scale"], scale_table).numpy(), g["gc_idx"])
    np.testing.assert_array_equal(O.gaussian_likelihood(y_hat, i["gc_scale"], i["gc_mu"]).numpy(), g["gc_lik"])
    zh, zl = O.entropy_bottleneck(i["eb_z"], sd)
    np.testing.assert_array_equal(zh.numpy(), g["eb_zhat"])
    np.testing.assert_allclose(zl.numpy(), g["eb_lik"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(O.window_attention(i["attn8_x"], sd, "g_a.4.conv_b.0", 8, 4).numpy(), g["attn8"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.window_attention(i["attn4_x"], sd, "g_a.8.conv_b.0", 4, 2).numpy(), g["attn4"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.gdn(i["gdn_x"], sd, "g_a.1").numpy(), g["gdn"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.gdn(i["gdn_x"], sd, "g_s.2", inverse=True).numpy(), g["igdn"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.deconv(i["deconv_x"], sd, "g_s.3").numpy(), g["deconv"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.residual_unit(i["block8_x"], sd, "g_a.4.conv_a.0").numpy(), g["ru"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.attention_block(i["block8_x"], sd, "g_a.4", 8, 4).numpy(), g["block8"], rtol=1e-5, atol=1e-5)


def test_index_build_equals_searchsorted(scale_table):
    """the 63-compare loop == searchsorted(table[:-1], max(s,.11), right=False)."""
    s = weights.hash_uniform("idx.s", (4096,)) ** 4 * 400.0 - 0.1
    s[:64] = scale_table
    ref = O.gc_indexes(s, scale_table)
    alt = torch.searchsorted(scale_table[:-1].contiguous(), torch.clamp(s, min=0.11), right=False).to(torch.int32)
    assert torch.equal(ref, alt)


def test_padding_rule():
    x = torch.ones(1, 3, 1365, 2048)
    xp, (left, top) = O.pad_to_multiple(x)
    assert xp.shape[-2:] == (1408, 2048) and (left, top) == (0, 21)
    assert xp[0, 0, 20].sum() == 0 and xp[0, 0, 21].sum() == 2048 and xp[0, 0, -22:].sum() == 0


def test_swin_block_oracle_matches_reference_tcm_block():
    """A14: the LN + W/SW-MSA + MLP block restatement vs the reference's own tcm.Block outputs."""
    from tests.golden.make_golden import SWIN_CASES, swin_state_dict
    g = _npz("ops.npz")
    for name, (dim, hd, ws, typ, B, H, W) in SWIN_CASES.items():
        sd = {f"blk.{k}": v for k, v in swin_state_dict(name, dim, hd, ws).items()}
        x = weights.hash_symmetric(f"{name}.x", (B, H, W, dim), 1.5)
        out = O.swin_block(x, sd, "blk", hd, ws, typ == "SW")
        np.testing.assert_allclose(out.numpy(), g[name], rtol=1e-5, atol=1e-5, err_msg=name)


def test_forward_matches_reference_lowrate_profile(scale_table):
    """Second operating point ("lowrate" weights: ~80 % zero symbols, bpp ~4): oracle vs the reference."""
    from tests.golden.make_golden import LOWRATE_CASES
    sd = weights.make_state_dict(seed=0, profile="lowrate")
    for case, (B, H, W) in LOWRATE_CASES.items():
        g = _npz(f"wacnn_{case}.npz")
        out = O.forward(sd, weights.make_image(B, H, W, seed=0), scale_table, collect=True)
        for key in ("y", "z", "latent_means", "latent_scales", "x_hat"):
            np.testing.assert_allclose(out[key].numpy(), g[key], rtol=1e-4, atol=1e-4, err_msg=key)
        assert (out["symbols"].numpy() != g["symbols"]).mean() <= 1e-3
        assert (g["symbols"] == 0).mean() > 0.7
