"""Golden fixture for the CDF-table build (`update()`).

Runs the UNMODIFIED reference `WACNN.update()` (cnn.py:135-140 -> GaussianConditional.update
entropy_models.py:599-625, EntropyBottleneck.update :356-394).  The only piece the reference
cannot execute here is `compressai._CXX.pmf_to_quantized_cdf` (C++ of the pip package, absent
from /root/reference): it is substituted by oracle/cdf_oracle.py's restatement of the published
algorithm, wrapped in a recorder.  The fixture therefore pins

  * every pmf row (float32, incl. the tail mass) exactly as the reference's Python computed it,
  * `_offset` and `_cdf_length` as the reference computed them,

and stores, for regression, the `_quantized_cdf` the restated integer stage produced from those
rows (that stage is "parity unpinned", see oracle/cdf_oracle.py).

    python tests/golden/make_golden_cdf.py      (build container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cdf_oracle, ref_shim, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sd = weights.make_state_dict(seed=0)
    net = ref_shim.reference_wacnn().eval()
    net.load_state_dict(sd, strict=True)
    em = sys.modules[type(net.entropy_bottleneck).__module__]
    recorded = []

    def recorder(pmf_list, precision):
        recorded.append(np.asarray(pmf_list, dtype=np.float32))
        assert all(float(np.float32(v)) == v for v in pmf_list)  # the list holds exact float32 values
        return cdf_oracle.pmf_to_quantized_cdf(pmf_list, precision).tolist()

    em._pmf_to_quantized_cdf = recorder
    assert net.update() is True
    gc, eb = net.gaussian_conditional, net.entropy_bottleneck
    n_gc, n_eb = gc.scale_table.numel(), eb.quantiles.shape[0]
    assert len(recorded) == n_gc + n_eb  # update_scale_table first, then the bottleneck (cnn.py:138-139)
    out = {}
    for name, mod, rows in (("gc", gc, recorded[:n_gc]), ("eb", eb, recorded[n_gc:])):
        width = max(r.size for r in rows)
        pad = np.zeros((len(rows), width), np.float32)
        for i, r in enumerate(rows):
            pad[i, : r.size] = r
        out[f"{name}_prob"] = pad                                   # pmf[:len] ++ tail_mass, zero padded
        out[f"{name}_prob_len"] = np.array([r.size for r in rows], np.int32)
        out[f"{name}_cdf"] = mod._quantized_cdf.numpy().astype(np.int32)
        out[f"{name}_cdf_length"] = mod._cdf_length.numpy().astype(np.int32)
        out[f"{name}_offset"] = mod._offset.numpy().astype(np.int32)
        assert np.array_equal(out[f"{name}_prob_len"] + 1, out[f"{name}_cdf_length"])
    out["scale_table"] = gc.scale_table.numpy()
    # the oracle's float stage against the reference's recorded rows
    for name, (pmf, tail, length, offset) in (("gc", cdf_oracle.gc_pmf(gc.scale_table)), ("eb", cdf_oracle.eb_pmf(sd))):
        for i in range(len(length)):
            row = np.concatenate([pmf[i, : int(length[i])].numpy(), tail[i].numpy()])
            np.testing.assert_allclose(row, out[f"{name}_prob"][i, : row.size], rtol=1e-6, atol=1e-12)
        assert np.array_equal(offset.numpy(), out[f"{name}_offset"])
    path = os.path.join(HERE, "cdf_tables.npz")
    np.savez_compressed(path, **out)
    print("gc rows", n_gc, "max len", out["gc_prob_len"].max(), "eb rows", n_eb, "max len", out["eb_prob_len"].max(),
          os.path.getsize(path))


if __name__ == "__main__":
    main()
