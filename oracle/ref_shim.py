"""TEST INFRASTRUCTURE ONLY -- import stubs that let the *unmodified* reference
python files under /root/reference be imported in the build container.

The reference (`src/compress/models/WACNN/cnn.py:5-7`) imports its entropy
models / GDN / ops from pip `compressai` and `timm`, neither of which is
installed here.  The reference tree vendors the same classes
(`src/compress/entropy_models/entropy_models.py`, `src/compress/layers/gdn.py`,
`src/compress/ops/*.py`), so we map the missing module names onto those
vendored files.  This file contains **no reference arithmetic**: every stub is
either inert (rANS, pmf_to_quantized_cdf: out of scope) or a one-line alias.

/root/reference does not exist on the GPU box.  Only
`tests/golden/make_golden.py` (run here, by hand) imports this module; nothing
in `-m gpu` tests, `smoke()` or `bench.py` does.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("RESDSIC_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "compress"))


def _new(name, is_pkg=False):
    m = types.ModuleType(name)
    if is_pkg:
        m.__path__ = []
    sys.modules[name] = m
    return m


def _exec(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class _Inert:
    """Placeholder for the absent rANS pybind classes (bitstream: out of scope)."""

    def __init__(self, *a, **k):
        pass


class _Identity(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x):
        return x


def install():
    """Idempotently install the stubs and put the reference on sys.path."""
    if "compressai.entropy_models" in sys.modules:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)

    # timm.models.layers: only DropPath(p=0) / to_2tuple / trunc_normal_ are used
    # (win_attention.py:3,81,147,150)
    timm = _new("timm", True)
    tmodels = _new("timm.models", True)
    tlayers = _new("timm.models.layers")
    tlayers.DropPath = _Identity
    tlayers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    tlayers.trunc_normal_ = lambda t, std=1.0, **kw: nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)
    timm.models = tmodels
    tmodels.layers = tlayers

    # metrics / plotting packages touched at import time by training helpers
    msssim = _new("pytorch_msssim")
    msssim.ms_ssim = lambda a, b, data_range=1.0: torch.zeros(())

    cai = _new("compressai", True)
    ops_pkg = _new("compressai.ops", True)
    bound = _exec("compressai.ops.bound_ops", f"{REF_SRC}/compress/ops/bound_ops.py")
    param = _exec("compressai.ops.parametrizers", f"{REF_SRC}/compress/ops/parametrizers.py")
    ops_pkg.LowerBound = bound.LowerBound
    ops_pkg.NonNegativeParametrizer = param.NonNegativeParametrizer

    def compute_padding(in_h, in_w, *, out_h=None, out_w=None, min_div=1):
        # rule of eval_model/__main__.py:89-95 (symmetric pad to a multiple)
        out_h = out_h if out_h is not None else -(-in_h // min_div) * min_div
        out_w = out_w if out_w is not None else -(-in_w // min_div) * min_div
        left = (out_w - in_w) // 2
        top = (out_h - in_h) // 2
        right, bottom = out_w - in_w - left, out_h - in_h - top
        return (left, right, top, bottom), (-left, -right, -top, -bottom)

    ops_pkg.compute_padding = compute_padding
    cai.ops = ops_pkg
    cai.available_entropy_coders = lambda: ["ans"]
    cai.get_entropy_coder = lambda: "ans"

    cxx = _new("compressai._CXX")

    def _no_cxx(*a, **k):
        raise NotImplementedError("compressai._CXX is not available (out of scope)")

    cxx.pmf_to_quantized_cdf = _no_cxx
    ans = _new("compressai.ans")
    ans.RansEncoder = ans.RansDecoder = ans.BufferedRansEncoder = _Inert
    cai.ans = ans
    _exec("compressai.entropy_models", f"{REF_SRC}/compress/entropy_models/entropy_models.py")
    _exec("compressai.layers", f"{REF_SRC}/compress/layers/gdn.py")


def reference_tcm_module():
    """Import the reference's (unregistered) `compress/models/TCM/tcm.py` -- the in-tree statement of the
    Swin LayerNorm + W/SW-MSA + MLP block (tcm.py:139-236).  Its remaining pip-compressai imports are only
    needed by the TCM *model* classes, so they are satisfied with inert placeholders."""
    install()
    layers = sys.modules["compressai.layers"]
    for name in ("AttentionBlock", "ResidualBlock", "ResidualBlockUpsample", "ResidualBlockWithStride"):
        if not hasattr(layers, name):
            setattr(layers, name, _Inert)
    from compress.layers.layers import conv3x3, subpel_conv3x3
    layers.conv3x3, layers.subpel_conv3x3 = conv3x3, subpel_conv3x3
    if "compressai.models" not in sys.modules:
        m = _new("compressai.models")
        m.CompressionModel = nn.Module
    tl = sys.modules["timm.models.layers"]
    tl.DropPath = _Identity
    import importlib
    return importlib.import_module("compress.models.TCM.tcm")


def reference_wacnn(N=192, M=320):
    """Instantiate the reference's own `WACNN` (cnn.py:23)."""
    install()
    from compress.models import models  # noqa: the reference registry

    return models["cnn"](N=N, M=M)
