"""TEST INFRASTRUCTURE -- CPU restatement of the CDF-table build behind `update()`
(SURVEY section 8f, N2).  Not product code: see oracle/__init__.py.

Two stages:

* float stage -- the pmf of every table row, restating the reference's own Python
  (`GaussianConditional.update` entropy_models/entropy_models.py:599-625,
  `EntropyBottleneck.update` :356-394, `_pmf_to_cdf` :174-182).  Pinned: the golden
  fixture tests/golden/cdf_tables.npz holds the pmf rows the UNMODIFIED reference code
  computed (tests/golden/make_golden_cdf.py).
* integer stage -- `pmf_to_quantized_cdf`.  In the reference this is C++ from the pip
  package `compressai` (`compressai._CXX`, version unpinned in the reference's Dockerfile:6;
  the fork calls itself 1.1.6dev0); its source is NOT in /root/reference and no reference
  test holds a known answer for it, so this stage is **parity unpinned**: the function
  below restates the published CompressAI algorithm (cpp_exts/ops/ops.cpp, itself a port of
  ryg_rans' frequency normalisation) and the CUDA kernel is held bit-exact against it.
"""
import math

import numpy as np
import torch

from . import wacnn_oracle as O


def pmf_to_quantized_cdf(pmf, precision=16):
    """CompressAI `pmf_to_quantized_cdf(pmf: List[float], precision) -> List[int]`.

    cdf[0] = 0, cdf[k+1] = round(p_k * 2^precision) (fp32 product, round half away from
    zero); rescale each entry by 2^precision / total (integer division); prefix sum; force
    the last entry to 2^precision; then, scanning left to right, give every zero-width
    symbol one count stolen from the lowest-frequency symbol that has more than one
    (first such symbol on ties), shifting the boundaries in between."""
    p = np.asarray(pmf, dtype=np.float32)
    if p.size == 0 or not np.all(np.isfinite(p)) or np.any(p < 0):
        raise ValueError("Invalid `pmf`, non-finite or negative element found")
    scaled = p * np.float32(1 << precision)                       # exact: power of two
    rounded = np.floor(scaled.astype(np.float64) + 0.5).astype(np.int64)   # std::round for values >= 0
    cdf = np.concatenate([[0], rounded]).astype(np.int64)
    total = int(cdf.sum())
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    cdf = ((1 << precision) * cdf) // total
    cdf = np.cumsum(cdf)
    cdf[-1] = 1 << precision
    n = cdf.size
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            freq = np.diff(cdf)
            cand = np.where(freq > 1, freq, np.iinfo(np.int64).max)
            best = int(np.argmin(cand))                            # first index of the minimum
            assert cand[best] != np.iinfo(np.int64).max, "no symbol to steal from"
            if best < i:
                cdf[best + 1:i + 1] -= 1
            else:
                assert best > i
                cdf[i + 1:best + 1] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision) and np.all(np.diff(cdf) > 0)
    return cdf.astype(np.int32)


def pmf_rows_to_cdf(pmf, tail_mass, pmf_length, max_length, precision=16):
    """EntropyModel._pmf_to_cdf (entropy_models.py:174-182): int32 [rows, max_length + 2], zero padded."""
    cdf = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
    for i in range(len(pmf_length)):
        prob = np.concatenate([np.asarray(pmf[i][: pmf_length[i]], np.float32), np.asarray(tail_mass[i], np.float32).reshape(1)])
        c = pmf_to_quantized_cdf(prob, precision)
        cdf[i, : c.size] = c
    return cdf


def standardized_quantile_multiplier(tail_mass=1e-9):
    """-scipy.stats.norm.ppf(tail_mass / 2) (entropy_models.py:587-588,601)."""
    import scipy.stats
    return float(-scipy.stats.norm.ppf(tail_mass / 2))


def gc_pmf(scale_table, tail_mass=1e-9):
    """GaussianConditional.update float stage (entropy_models.py:599-619).
    Returns pmf [T, max_length], tail_mass [T,1], pmf_length [T], offset [T]."""
    table = torch.as_tensor(scale_table, dtype=torch.float32)
    multiplier = standardized_quantile_multiplier(tail_mass)
    pmf_center = torch.ceil(table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(pmf_length.max())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    scale = table.unsqueeze(1)
    c = float(-(2 ** -0.5))
    upper = 0.5 * torch.erfc(c * ((0.5 - samples) / scale))
    lower = 0.5 * torch.erfc(c * ((-0.5 - samples) / scale))
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    return pmf, tail, pmf_length, -pmf_center


def eb_pmf(sd, p="entropy_bottleneck"):
    """EntropyBottleneck.update float stage (entropy_models.py:356-390).
    Returns pmf [C, max_length], tail_mass [C,1], pmf_length [C], offset [C]."""
    q = sd[p + ".quantiles"]
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    lower = O.eb_logits_cumulative(samples - 0.5, sd, p)
    upper = O.eb_logits_cumulative(samples + 0.5, sd, p)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    return pmf, tail, pmf_length, -minima


def gc_update(scale_table, tail_mass=1e-9, precision=16):
    pmf, tail, length, offset = gc_pmf(scale_table, tail_mass)
    cdf = pmf_rows_to_cdf(pmf.numpy(), tail.numpy(), length.numpy(), int(length.max()), precision)
    return {"quantized_cdf": cdf, "cdf_length": (length + 2).numpy().astype(np.int32), "offset": offset.numpy().astype(np.int32)}


def eb_update(sd, p="entropy_bottleneck", precision=16):
    pmf, tail, length, offset = eb_pmf(sd, p)
    cdf = pmf_rows_to_cdf(pmf.numpy(), tail.numpy(), length.numpy(), int(length.max()), precision)
    return {"quantized_cdf": cdf, "cdf_length": (length + 2).numpy().astype(np.int32), "offset": offset.numpy().astype(np.int32)}
