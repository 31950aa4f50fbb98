"""`-m stf` (builder-defined: the reference has no STF model).  Model-level parity is against
oracle/stf_oracle.py -- a PyTorch composition of blocks that ARE pinned to the reference (Swin block =
tcm.Block, entropy models / slice loop = cnn).  CPU part: host-side program via the descriptor interpreter."""
import numpy as np
import pytest
import torch

import resdsic_b200
from oracle import stf_oracle as S
from oracle import weights
from tests.helpers import compare_forward
from tests.program_sim import run_on_cpu


@pytest.fixture(scope="module")
def stf():
    m = resdsic_b200.models["stf"]().eval()
    sd = weights.synth_state_dict(m.state_dict())
    m.load_state_dict(sd, strict=True)
    return m, sd


def _as_golden(ref):
    return dict(y=ref["y"].numpy(), z=ref["z"].numpy(), latent_means=ref["latent_means"].numpy(),
                latent_scales=ref["latent_scales"].numpy(), y_hat=ref["y_hat"].numpy(), x_hat=ref["x_hat"].numpy(),
                lik_y=ref["likelihoods"]["y"].numpy(), lik_z=ref["likelihoods"]["z"].numpy(),
                symbols=ref["symbols"].numpy(), indexes=ref["indexes"].numpy())


def _collect(p, M):
    return dict(y=p.y.to_nchw().float().cpu().numpy(), z=p.z.to_nchw().float().cpu().numpy(),
                latent_means=p.means.channels(0, M).to_nchw().float().cpu().numpy(),
                latent_scales=p.scales.channels(0, M).to_nchw().float().cpu().numpy(),
                y_hat=p.y_hat.to_nchw().cpu().numpy(), x_hat=p.x_hat.cpu().numpy(), lik_y=p.lik_y.cpu().numpy(),
                lik_z=p.lik_z.cpu().numpy(), symbols=p.symbols.cpu().numpy(), indexes=p.indexes.cpu().numpy())


def test_registry_has_stf(stf):
    m, sd = stf
    assert resdsic_b200.models["stf"] is resdsic_b200.SymmetricalTransFormer

    class A:
        model = "stf"
    assert isinstance(resdsic_b200.configure_model(A), resdsic_b200.SymmetricalTransFormer)
    assert m.num_slices == 12 and m.max_support_slices == 6 and m.M == 384
    assert [n for n, _ in m.named_parameters() if n.endswith(".quantiles")] == ["entropy_bottleneck.quantiles"]


@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (2, 128, 192)])
def test_stf_program_matches_oracle_cpu(stf, B, H, W):
    m, sd = stf
    table = weights.scale_table()
    x = weights.make_image(B, H, W, seed=9)
    ref = S.forward(sd, x, table, collect=True)
    p = m._build(B, H, W, "cpu", True, build_only=True)
    p.x.copy_(x)
    run_on_cpu(p.prog)
    # y spans +-14 here, so a 1e-5 wobble crosses more round-half ties than in the cnn fixtures
    compare_forward(_collect(p, m.M), _as_golden(ref), B * H * W, cont_tol=2e-4, flip_frac=5e-2, yhat_frac=1.0,
                    xhat_max=0.1, xhat_psnr=35.0)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_stf_forward_gpu_vs_oracle(stf, precision):
    m, sd = stf
    dev = "cuda:0"
    m = m.to(dev).set_precision(precision)
    table = weights.scale_table()
    x = weights.make_image(2, 128, 192, seed=9)
    ref = _as_golden(S.forward(sd, x, table, collect=True))
    r = m.symbols_and_indexes(x.to(dev))
    p = m._last_plan
    got = _collect(p, m.M)
    assert r["x_hat"].shape == (2, 3, 128, 192) and r["likelihoods"]["y"].shape == (2, 384, 8, 12)
    if precision == "fp32":
        stats = compare_forward(got, ref, 2 * 128 * 192, cont_tol=1e-3, flip_frac=5e-2, yhat_frac=1.0, xhat_max=0.2,
                                xhat_psnr=30.0)
    else:
        dy = np.abs(got["y"] - ref["y"])
        flips = (got["symbols"] != ref["symbols"]).mean()
        from tests.helpers import bpp_of
        n = 2 * 128 * 192
        b_got, b_ref = bpp_of(got["lik_y"], got["lik_z"], n), bpp_of(ref["lik_y"], ref["lik_z"], n)
        stats = dict(y_max=float(dy.max()), y_mean=float(dy.mean()), flips=float(flips), bpp=b_got, bpp_ref=b_ref)
        assert dy.mean() <= 0.08 and flips <= 0.25 and abs(b_got - b_ref) <= 3e-2 * b_ref
    print("stf", precision, stats)
    m.set_precision("fp32").cpu()
