"""Decoder-side slice loop (reference WACNN.decompress, cnn.py:296-342, with the two entropy-coder
calls left to the caller).  Golden: tests/golden/wacnn_decode_c64x128.npz -- the UNMODIFIED reference decompress
run with replay stubs for `entropy_bottleneck.decompress` / `RansDecoder` (tests/golden/make_golden_decode.py)."""
import os

import numpy as np
import pytest
import torch

import resdsic_b200
from oracle import wacnn_oracle as O
from oracle import weights
from tests.conftest import GOLDEN
from tests.golden.make_golden_decode import CASES
from tests.program_sim import run_on_cpu

CASE = "decode_c64x128"
B, H, W = CASES[CASE]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, f"wacnn_{CASE}.npz"))


def test_oracle_decode_matches_reference(gold, synthetic_sd, scale_table):
    out = O.decode(synthetic_sd, torch.from_numpy(gold["z_hat"]), torch.from_numpy(gold["symbols"]), scale_table)
    np.testing.assert_array_equal(out["indexes"].numpy(), gold["indexes"])
    np.testing.assert_allclose(out["x_hat"].numpy(), gold["x_hat_dec"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["y_hat"].numpy(), gold["y_hat"], rtol=0, atol=1e-4)
    # what the reference guarantees and the codec relies on: decoder output == clamp(encoder-side reconstruction)
    np.testing.assert_array_equal(gold["x_hat_dec"], np.clip(gold["x_hat_enc"], 0, 1))


def test_decoder_programs_reproduce_reference(gold, synthetic_sd):
    """The host-side decoder plan (22 programs) interpreted on CPU."""
    m = resdsic_b200.WACNN().eval()
    m.load_state_dict(synthetic_sd, strict=True)
    p = m._build_decoder(B, H // 64, W // 64, "cpu", build_only=True)
    assert len(p.params) == 10 and len(p.update) == 10
    p.z_hat_in.copy_(torch.from_numpy(gold["z_hat"]))
    run_on_cpu(p.hyper)
    sym = torch.from_numpy(gold["symbols"])
    for i in range(10):
        run_on_cpu(p.params[i])
        frac = (p.indexes[:, 32 * i:32 * i + 32].numpy() != gold["indexes"][:, 32 * i:32 * i + 32]).mean()
        assert frac <= 1e-3, (i, frac)
        p.symbols[:, 32 * i:32 * i + 32].copy_(sym[:, 32 * i:32 * i + 32])
        run_on_cpu(p.update[i])
    run_on_cpu(p.synth)
    assert np.abs(p.x_hat.numpy() - gold["x_hat_dec"]).max() <= 1e-3
    assert p.x_hat.min() >= 0 and p.x_hat.max() <= 1


@pytest.fixture(scope="module")
def model(synthetic_sd):
    return resdsic_b200.WACNN.from_state_dict(synthetic_sd).to("cuda:0").eval()


@pytest.mark.gpu
def test_slice_decoder_fp32_vs_reference_golden(model, gold):
    model.set_precision("fp32")
    dec = model.slice_decoder(torch.from_numpy(gold["z_hat"]).cuda())
    sym = torch.from_numpy(gold["symbols"]).cuda()
    with pytest.raises(RuntimeError, match="sequential"):
        dec.indexes(3)
    with pytest.raises(RuntimeError, match="first"):
        dec.push_symbols(0, sym[:, :32])
    for i in range(10):
        idx = dec.indexes(i)
        frac = (idx.cpu().numpy() != gold["indexes"][:, 32 * i:32 * i + 32]).mean()
        assert frac <= 1e-3, (i, frac)
        dec.push_symbols(i, sym[:, 32 * i:32 * i + 32])
    x = dec.finish()
    assert np.abs(x.cpu().numpy() - gold["x_hat_dec"]).max() <= 1e-3
    assert np.abs(dec.y_hat.cpu().numpy() - gold["y_hat"]).max() <= 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("graph", [True, False])
def test_decoder_is_bit_identical_to_encoder_pass(model, precision, graph):
    """Round trip without the entropy coder: the decoder, fed the encoder's symbols, asks for EXACTLY the CDF
    indexes the encoder used (or rANS would desynchronise) and reconstructs EXACTLY clamp(x_hat) of the forward."""
    model.set_precision(precision)
    model.use_cuda_graph = graph
    try:
        x = weights.make_image(3, 128, 192, seed=9).cuda()
        r = model.symbols_and_indexes(x)
        ysym, yidx, zsym = r["y_symbols"].clone(), r["y_indexes"].clone(), r["z_symbols"].clone()
        x_hat = r["x_hat"].clone()
        med = model.entropy_bottleneck._get_medians().detach().view(1, -1, 1, 1)
        z_hat = model.entropy_bottleneck.dequantize(zsym, med)      # what entropy_bottleneck.decompress returns
        dec = model.slice_decoder(z_hat)
        for i in range(model.num_slices):
            assert torch.equal(dec.indexes(i), yidx[:, 32 * i:32 * i + 32]), i
            dec.push_symbols(i, ysym[:, 32 * i:32 * i + 32])
        assert torch.equal(dec.finish(), x_hat.clamp(0, 1))
        with pytest.raises(RuntimeError):
            model.slice_decoder(z_hat).finish()
    finally:
        model.set_precision("fp32")
        model.use_cuda_graph = True


def test_bf16_encoder_and_decoder_programs_agree_bit_for_bit_on_the_cpu_simulator(synthetic_sd):
    """bf16 mode, no GPU: the encoder-side program (grouped slice loop, grouped hyper-synthesis convs, grouped tail with the
    three-way LRP split) and the decoder plan (one launch per convolution, mirrored LRP split) produce the same CDF index
    for every latent and, with the encoder's symbols pushed back, the same x_hat -- what the entropy decoder relies on.
    The simulator evaluates a group as its own convolution, as the GPU kernel evaluates it as its own n tile."""
    m = resdsic_b200.WACNN().eval()
    m.load_state_dict(synthetic_sd, strict=True)
    m.set_precision("bf16")
    Bq, Hq, Wq = 2, 128, 192
    enc = m._build(Bq, Hq, Wq, "cpu", True, build_only=True)
    assert enc.prog.num_launches < 200
    enc.x.copy_(weights.make_image(Bq, Hq, Wq, seed=3))
    run_on_cpu(enc.prog)
    med = m.entropy_bottleneck._get_medians().detach().view(1, -1, 1, 1)
    z_hat = enc.z_symbols.float() + med  # what EntropyBottleneck.decompress hands the decoder (entropy_models.py:521-526)
    dec = m._build_decoder(Bq, Hq // 64, Wq // 64, "cpu", build_only=True)
    dec.z_hat_in.copy_(z_hat)
    run_on_cpu(dec.hyper)
    for i in range(10):
        sl = slice(32 * i, 32 * i + 32)
        run_on_cpu(dec.params[i])
        assert torch.equal(dec.indexes[:, sl], enc.indexes[:, sl]), (i, (dec.indexes[:, sl] != enc.indexes[:, sl]).float().mean())
        dec.symbols[:, sl].copy_(enc.symbols[:, sl])
        run_on_cpu(dec.update[i])
    run_on_cpu(dec.synth)
    assert torch.equal(dec.x_hat, enc.x_hat.clamp(0, 1))
