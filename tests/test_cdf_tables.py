"""CDF-table build (`update()`): GaussianConditional.update (reference
entropy_models.py:599-625), EntropyBottleneck.update (:356-394), `_pmf_to_cdf` (:174-182) and
`pmf_to_quantized_cdf` (pip compressai C++, restated -- see oracle/cdf_oracle.py for what is and is
not pinned).  Golden: tests/golden/cdf_tables.npz (tests/golden/make_golden_cdf.py)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import resdsic_b200
from oracle import cdf_oracle as CO
from tests.conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "cdf_tables.npz"))


# ------------------------------------------------------------------ CPU: the oracle
def test_pmf_to_quantized_cdf_known_answers():
    np.testing.assert_array_equal(CO.pmf_to_quantized_cdf([0.5, 0.25, 0.25]), [0, 32768, 49152, 65536])
    # zero-width symbols on both sides of the only populated one: steal right (i < donor), then left (i > donor)
    np.testing.assert_array_equal(CO.pmf_to_quantized_cdf([0.0, 1.0, 0.0], precision=4), [0, 1, 15, 16])
    # ties: the FIRST lowest-frequency donor with more than one count gives
    np.testing.assert_array_equal(CO.pmf_to_quantized_cdf([0.25, 0.25, 0.0, 0.25, 0.25], precision=3), [0, 1, 3, 4, 6, 8])
    with pytest.raises(ValueError):
        CO.pmf_to_quantized_cdf([0.0, 0.0])
    with pytest.raises(ValueError):
        CO.pmf_to_quantized_cdf([0.5, -0.1])


def test_oracle_float_stage_matches_reference_rows(gold, synthetic_sd, scale_table):
    for name, (pmf, tail, length, offset) in (("gc", CO.gc_pmf(scale_table)), ("eb", CO.eb_pmf(synthetic_sd))):
        assert np.array_equal((length + 2).numpy(), gold[f"{name}_cdf_length"])
        assert np.array_equal(offset.numpy(), gold[f"{name}_offset"])
        for i in range(len(length)):
            row = np.concatenate([pmf[i, : int(length[i])].numpy(), tail[i].numpy()])
            np.testing.assert_allclose(row, gold[f"{name}_prob"][i, : row.size], rtol=1e-6, atol=1e-12)


def test_oracle_tables_regression_and_invariants(gold):
    for name in ("gc", "eb"):
        cdf, n = gold[f"{name}_cdf"], gold[f"{name}_cdf_length"]
        for i in (0, 1, len(n) // 2, len(n) - 1):
            row = CO.pmf_to_quantized_cdf(gold[f"{name}_prob"][i, : n[i] - 1])
            np.testing.assert_array_equal(row, cdf[i, : n[i]])
        for i in range(len(n)):
            r = cdf[i, : n[i]]
            assert r[0] == 0 and r[-1] == 65536 and (np.diff(r) > 0).all() and (cdf[i, n[i]:] == 0).all()


# ------------------------------------------------------------------ GPU
def _dev_cdf(prob, cdf_length, precision=16):
    from resdsic_b200 import _lib
    L = _lib.lib()
    p = torch.as_tensor(prob, dtype=torch.float32).cuda().contiguous()
    n = torch.as_tensor(cdf_length, dtype=torch.int32).cuda()
    cdf = torch.full((p.shape[0], p.shape[1] + 1), -1, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = L.rdsic_pmf_to_quantized_cdf(p.data_ptr(), p.shape[1], n.data_ptr(), p.shape[0], precision, cdf.data_ptr(),
                                      cdf.shape[1], status.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    return cdf.cpu().numpy(), int(status.item())


@pytest.mark.gpu
def test_integer_stage_bit_exact_on_reference_rows(gold):
    for name in ("gc", "eb"):
        got, status = _dev_cdf(gold[f"{name}_prob"], gold[f"{name}_cdf_length"])
        assert status == 0
        np.testing.assert_array_equal(got, gold[f"{name}_cdf"], err_msg=name)


@pytest.mark.gpu
def test_integer_stage_bit_exact_on_adversarial_rows():
    rng = np.random.default_rng(5)
    rows, width = 96, 300
    prob = np.zeros((rows, width), np.float32)
    n = np.zeros(rows, np.int32)
    for r in range(rows):
        k = int(rng.integers(2, width + 1))
        v = rng.random(k).astype(np.float32) ** int(rng.integers(1, 12))     # heavy-tailed: many entries round to 0
        v[rng.random(k) < rng.random() * 0.8] = 0.0
        v[int(rng.integers(0, k))] += 1.0                                     # at least one donor
        prob[r, :k] = v / v.sum() * np.float32(rng.uniform(0.5, 1.5))          # totals away from 2^16 exercise the rescale
        n[r] = k + 1
    for precision in (16, 12):
        got, status = _dev_cdf(prob, n, precision)
        assert status == 0
        for r in range(rows):
            want = CO.pmf_to_quantized_cdf(prob[r, : n[r] - 1], precision)
            np.testing.assert_array_equal(got[r, : n[r]], want, err_msg=f"row {r} precision {precision}")
            assert (got[r, n[r]:] == 0).all()
    # ragged edge: a single probability
    got, status = _dev_cdf(np.array([[0.3, 0.0]], np.float32), [2])
    np.testing.assert_array_equal(got[0, :2], [0, 65536])
    # invalid rows are reported, not silently normalised
    bad = prob[:3].copy()
    bad[1, 0] = -0.25
    assert _dev_cdf(bad, n[:3])[1] == 2
    bad[1, :] = 0.0
    assert _dev_cdf(bad, n[:3])[1] == 2


@pytest.fixture(scope="module")
def model(synthetic_sd):
    return resdsic_b200.WACNN.from_state_dict(synthetic_sd).to("cuda:0").eval()


@pytest.mark.gpu
def test_model_update_vs_reference_golden(model, gold, scale_table):
    gc, eb = model.gaussian_conditional, model.entropy_bottleneck
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        gc._check_cdf_size()
    assert model.update() is True
    assert torch.equal(gc.scale_table.cpu(), scale_table)
    for name, mod in (("gc", gc), ("eb", eb)):
        mod._check_cdf_size(), mod._check_offsets_size(), mod._check_cdf_length()
        assert mod._quantized_cdf.dtype == torch.int32 and mod._offset.dtype == torch.int32
        np.testing.assert_array_equal(mod._offset.cpu().numpy(), gold[f"{name}_offset"])
        np.testing.assert_array_equal(mod._cdf_length.cpu().numpy(), gold[f"{name}_cdf_length"])
        n = gold[f"{name}_cdf_length"]
        prob = mod._last_prob.cpu().numpy()
        assert prob.shape == gold[f"{name}_prob"].shape
        # float stage: CUDA erfcf / expf / tanhf vs the reference's libm
        np.testing.assert_allclose(prob, gold[f"{name}_prob"], rtol=2e-4, atol=2e-9, err_msg=name)
        cdf = mod._quantized_cdf.cpu().numpy()
        assert cdf.shape == gold[f"{name}_cdf"].shape
        # integer stage on the device's own rows: bit-exact against the oracle
        for i in range(len(n)):
            np.testing.assert_array_equal(cdf[i, : n[i]], CO.pmf_to_quantized_cdf(prob[i, : n[i] - 1]), err_msg=f"{name} row {i}")
        # end to end against the tables built from the reference's rows, compared as symbol FREQUENCIES (the
        # cumulative form carries every one-count difference forward): an ulp-level difference between CUDA erfcf
        # and libm erfc moves round(p * 65536) by one count where p * 65536 sits next to a .5 boundary (p * 65536
        # reaches ~100 in the wide rows, so a 2e-4 relative error flips a few percent of the symbols), never more
        # than that -- encoder and decoder must therefore build their tables with the same implementation, as with
        # CompressAI itself across platforms.
        # A flipped count also changes `total`, hence the integer rescale of EVERY symbol by at most one, and the
        # last symbol (the tail mass) absorbs the remainder, and the donor of the stolen counts may be a different
        # symbol: the meaningful end-to-end bound is the total variation between the two quantised pmfs (< 1 %).
        for i in range(len(n)):
            fd = np.abs(np.diff(cdf[i, : n[i]].astype(np.int64)) - np.diff(gold[f"{name}_cdf"][i, : n[i]].astype(np.int64)))
            assert fd.sum() <= 0.01 * 65536, (name, i, fd.max(), fd.sum())
    # the forward path keeps working with the buffers filled (scale table now comes from the buffer)
    from oracle import weights
    out = model(weights.make_image(1, 64, 64, seed=0).cuda())
    assert torch.isfinite(out["x_hat"]).all()
