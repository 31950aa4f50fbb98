"""Host-side logic on CPU: the program the WACNN module builds (packing, phase
mapping, slot plan, addressing) interpreted by tests/program_sim.py must
reproduce the reference's golden outputs.  No CUDA compute here."""
import os

import numpy as np
import pytest
import torch

import resdsic_b200
from resdsic_b200.layers import Ctx
from tests.conftest import GOLDEN
from tests.golden.make_golden import CASES, op_inputs
from tests.helpers import compare_forward
from tests.program_sim import run_on_cpu
from oracle import weights


@pytest.fixture(scope="module")
def model(synthetic_sd):
    m = resdsic_b200.WACNN().eval()
    missing = m.load_state_dict(synthetic_sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def test_registry_and_state_dict_contract(model, synthetic_sd):
    assert "cnn" in resdsic_b200.models and resdsic_b200.models["cnn"] is resdsic_b200.WACNN
    sd = model.state_dict()
    assert list(sd.keys()) == list(synthetic_sd.keys()) and len(sd) == 585
    assert [n for n, _ in model.named_parameters() if n.endswith(".quantiles")] == ["entropy_bottleneck.quantiles"]
    assert model.num_slices == 10 and model.max_support_slices == 5

    class A:
        model, N, M = "cnn", 192, 320
    assert isinstance(resdsic_b200.configure_model(A), resdsic_b200.WACNN)


def test_no_cpu_fallback(model):
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError, match="multiple of 64"):
        model._plan(1, 100, 64, "cpu", False)


@pytest.mark.parametrize("case", list(CASES))
def test_program_reproduces_reference(case, model):
    B, H, W = CASES[case]
    g = np.load(os.path.join(GOLDEN, f"wacnn_{case}.npz"))
    p = model._build(B, H, W, "cpu", True, build_only=True)
    p.x.copy_(weights.make_image(B, H, W, seed=0))
    run_on_cpu(p.prog)
    assert p.prog.num_launches < 400
    got = dict(y=p.y.to_nchw().numpy(), z=p.z.to_nchw().numpy(),
               latent_means=p.means.channels(0, 320).to_nchw().numpy(),
               latent_scales=p.scales.channels(0, 320).to_nchw().numpy(), y_hat=p.y_hat.to_nchw().numpy(),
               x_hat=p.x_hat.numpy(), lik_y=p.lik_y.numpy(), lik_z=p.lik_z.numpy(),
               symbols=p.symbols.numpy(), indexes=p.indexes.numpy())
    compare_forward(got, g, B * H * W)
    # support slices landed in both context buffers
    np.testing.assert_array_equal(p.means.channels(320, 160).to_nchw().numpy(), p.scales.channels(320, 160).to_nchw().numpy())
    np.testing.assert_array_equal(p.means.channels(320, 160).to_nchw().numpy(), p.y_hat.channels(0, 160).to_nchw().numpy())


def test_layer_programs_reproduce_reference(model):
    g, i = np.load(os.path.join(GOLDEN, "ops.npz")), op_inputs()

    def run(module, x):
        ctx = Ctx("cpu", "fp32", build_only=True)
        out = ctx.to_nchw(module.emit(ctx, ctx.from_nchw(x)))
        run_on_cpu(ctx.prog)
        return out.numpy()

    np.testing.assert_allclose(run(model.g_a[4].conv_b[0], i["attn8_x"]), g["attn8"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(run(model.g_a[8].conv_b[0], i["attn4_x"]), g["attn4"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(run(model.g_a[1], i["gdn_x"]), g["gdn"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(run(model.g_s[2], i["gdn_x"]), g["igdn"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(run(model.g_s[3], i["deconv_x"]), g["deconv"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(run(model.g_a[4], i["block8_x"]), g["block8"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(run(model.g_a[4].conv_a[0], i["block8_x"]), g["ru"], rtol=1e-5, atol=1e-5)


def test_swin_block_program_reproduces_reference():
    """A14 (stf): the B200 Swin block's program (LN, qkv, window attention, proj+res, LN, MLP) interpreted on
    CPU vs the reference's tcm.Block goldens; its state_dict keys are the reference's."""
    from resdsic_b200.layers import SwinBlock
    from tests.golden.make_golden import SWIN_CASES, swin_state_dict
    g = np.load(os.path.join(GOLDEN, "ops.npz"))
    for name, (dim, hd, ws, typ, B, H, W) in SWIN_CASES.items():
        blk = SwinBlock(dim, dim, hd, ws, 0.0, typ).eval()
        res = blk.load_state_dict(swin_state_dict(name, dim, hd, ws), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        x = weights.hash_symmetric(f"{name}.x", (B, H, W, dim), 1.5)
        ctx = Ctx("cpu", "fp32", build_only=True)
        out = ctx.to_nchw(blk.emit(ctx, ctx.from_nchw(x.permute(0, 3, 1, 2).contiguous())))
        run_on_cpu(ctx.prog)
        assert ctx.prog.num_launches == 9  # 2 layout copies + 7 kernels
        np.testing.assert_allclose(out.permute(0, 2, 3, 1).numpy(), g[name], rtol=1e-4, atol=1e-4, err_msg=name)
