"""Training-mode forward VALUES (BASELINE config 4, forward half): noise quantisation in
EntropyBottleneck / GaussianConditional (reference entropy_models.py:131-137,447-490,646-661),
WACNN.forward in `.train()` mode (cnn.py:143-193) and aux_loss (WACNN/base.py:22-27).

Golden: tests/golden/wacnn_train_c64x128.npz, produced by the UNMODIFIED reference in train() mode
(tests/golden/make_golden_train.py), with the reference's own noise draws stored alongside.
Forward values only; gradients are covered by tests/test_training_step.py."""
import os

import numpy as np
import pytest
import torch

import resdsic_b200
from oracle import wacnn_oracle as O
from oracle import weights
from tests.conftest import GOLDEN
from tests.golden.make_golden_train import CASES
from tests.helpers import bpp_of
from tests.program_sim import run_on_cpu

CASE = "train_c64x128"
B, H, W = CASES[CASE]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, f"wacnn_{CASE}.npz"))


def _noise(gold):
    return {"y": torch.from_numpy(gold["noise_y"]), "z": torch.from_numpy(gold["noise_z"])}


def _check(got_xhat, got_ly, got_lz, gold, cont_tol):
    np.testing.assert_allclose(got_lz, gold["lik_z"], rtol=1e-3, atol=1e-8)
    bad = np.abs(got_ly - gold["lik_y"]) > 1e-4 + 1e-3 * gold["lik_y"]
    assert bad.mean() <= 1e-2, bad.mean()
    b_got, b_ref = bpp_of(got_ly, got_lz, B * H * W), bpp_of(gold["lik_y"], gold["lik_z"], B * H * W)
    assert abs(b_got - b_ref) <= 1e-3 * b_ref, (b_got, b_ref)
    assert np.abs(got_xhat - gold["x_hat"]).max() <= cont_tol


# ------------------------------------------------------------------ CPU: oracle and host program
def test_oracle_training_forward_matches_reference(gold, synthetic_sd):
    x = weights.make_image(B, H, W, seed=int(gold["image_seed"]))
    out = O.forward(synthetic_sd, x, noise=_noise(gold))
    np.testing.assert_allclose(out["likelihoods"]["y"].numpy(), gold["lik_y"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["likelihoods"]["z"].numpy(), gold["lik_z"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(out["x_hat"].numpy(), gold["x_hat"], rtol=0, atol=1e-4)
    assert abs(O.eb_aux_loss(synthetic_sd).item() - float(gold["aux_loss"])) <= 1e-5 * float(gold["aux_loss"])
    # the noise draw really is what separates the two modes
    ev = O.forward(synthetic_sd, x)
    assert np.abs(ev["likelihoods"]["y"].numpy() - gold["lik_y"]).max() > 1e-2


def test_training_program_reproduces_reference(gold, synthetic_sd):
    """Host-side builder in train() mode (noise buffers wired into the EB / GC descriptors), interpreted on CPU."""
    m = resdsic_b200.WACNN().train()
    m.load_state_dict(synthetic_sd, strict=True)
    p = m._build(B, H, W, "cpu", False, build_only=True)
    assert p.noise_y is not None and p.noise_z is not None
    p.x.copy_(weights.make_image(B, H, W, seed=int(gold["image_seed"])))
    nz = _noise(gold)
    p.noise_y.t.view(B, H // 16, W // 16, 320).copy_(nz["y"].permute(0, 2, 3, 1))
    p.noise_z.t.view(B, H // 64, W // 64, 192).copy_(nz["z"].permute(0, 2, 3, 1))
    run_on_cpu(p.prog)
    _check(p.x_hat.numpy(), p.lik_y.numpy(), p.lik_z.numpy(), gold, 1e-3)
    ev = resdsic_b200.WACNN().eval()
    assert ev._build(1, 64, 64, "cpu", False, build_only=True).noise_y is None


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def model(synthetic_sd):
    m = resdsic_b200.WACNN.from_state_dict(synthetic_sd).to("cuda:0")
    return m


@pytest.mark.gpu
def test_training_forward_fp32_vs_reference_golden(model, gold):
    model.set_precision("fp32").train()
    try:
        model.noise_override = _noise(gold)
        x = weights.make_image(B, H, W, seed=int(gold["image_seed"])).to("cuda:0")
        out = model(x)  # autograd enabled: the differentiable training forward (resdsic_b200/training)
        assert out["x_hat"].requires_grad and out["likelihoods"]["y"].requires_grad
        _check(out["x_hat"].detach().cpu().numpy(), out["likelihoods"]["y"].detach().cpu().numpy(),
               out["likelihoods"]["z"].detach().cpu().numpy(), gold, 1e-3)
        with torch.no_grad():  # no autograd: the planned program with the noise views (forward values only)
            out = model(x)
        assert not out["x_hat"].requires_grad
        _check(out["x_hat"].cpu().numpy(), out["likelihoods"]["y"].cpu().numpy(), out["likelihoods"]["z"].cpu().numpy(),
               gold, 1e-3)
        # own device-side draw: a different noise sample each call, same x_hat (ste_round path is noise-free)
        model.noise_override = None
        with torch.no_grad():
            a = {k: v.clone() for k, v in model(x)["likelihoods"].items()}
            xa = model(x)["x_hat"].clone()
            b = model(x)
        assert not torch.equal(a["y"], b["likelihoods"]["y"]) and not torch.equal(a["z"], b["likelihoods"]["z"])
        assert torch.equal(xa, b["x_hat"])
        assert (b["likelihoods"]["y"] > 0).all() and (b["likelihoods"]["y"] <= 1).all()
    finally:
        model.noise_override = None
        model.eval()


@pytest.mark.gpu
def test_aux_loss_vs_reference_golden(model, gold):
    got = float(model.aux_loss())
    assert abs(got - float(gold["aux_loss"])) <= 1e-5 * float(gold["aux_loss"]), (got, float(gold["aux_loss"]))


@pytest.mark.gpu
def test_entropy_modules_training_mode_vs_oracle(model, synthetic_sd):
    """Standalone modules with training=True: outputs = inputs + noise exactly, likelihood at outputs."""
    hs = weights.hash_symmetric
    gc, eb = model.gaussian_conditional, model.entropy_bottleneck
    y, mu = hs("tr.gc.y", (2, 32, 8, 12), 12.0), hs("tr.gc.mu", (2, 32, 8, 12), 3.0)
    sc = weights.hash_uniform("tr.gc.scale", (2, 32, 8, 12)) ** 4 * 300.0 - 0.05
    n = weights.hash_uniform("tr.gc.noise", (2, 32, 8, 12)) - 0.5
    out, lik = gc(y.cuda(), sc.cuda(), mu.cuda(), training=True, noise=n.cuda())
    assert torch.equal(out.cpu(), y + n)
    ref = O.gaussian_likelihood(y + n, sc, mu)
    np.testing.assert_allclose(lik.cpu().numpy(), ref.numpy(), rtol=3e-4, atol=1e-9)
    assert torch.equal(gc.quantize(y.cuda(), "noise", mu.cuda(), noise=n.cuda()).cpu(), y + n)
    drawn = gc.quantize(y.cuda(), "noise", mu.cuda()).cpu() - y
    assert drawn.abs().max() <= 0.5 + 1e-6 and drawn.std() > 0.2
    z = hs("tr.eb.z", (2, 192, 3, 5), 6.0)
    nz = weights.hash_uniform("tr.eb.noise", (2, 192, 3, 5)) - 0.5
    out, lik = eb(z.cuda(), training=True, noise=nz.cuda())
    assert torch.equal(out.cpu(), z + nz)
    _, ref = O.entropy_bottleneck(z, synthetic_sd, noise=nz)
    np.testing.assert_allclose(lik.cpu().numpy(), ref.numpy(), rtol=3e-4, atol=1e-9)
    # eval mode of the same modules is untouched
    out_e, _ = eb(z.cuda())
    zh, _ = O.entropy_bottleneck(z, synthetic_sd)
    assert torch.equal(out_e.cpu(), zh)
