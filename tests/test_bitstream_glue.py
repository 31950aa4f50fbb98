"""compress() / decompress() (reference cnn.py:217-342, entropy_models.py:205-287,512-526): everything around the
rANS coder.  The coder itself (`compressai.ans`, C++ of pip CompressAI) is out of scope and absent from this image,
so the tests install a stand-in module with the same call surface whose "bitstream" simply carries the symbols AND the
CDF indexes they were coded with -- which lets the decoder side assert that it asks for exactly the encoder's
indexes, in the reference's order, with arguments of the types the real pybind coder requires."""
import pickle
import sys
import types

import pytest
import torch

import resdsic_b200
from oracle import weights


def _check_args(symbols, indexes, cdf, cdf_lengths, offsets):
    assert isinstance(indexes, list) and all(type(v) is int for v in indexes[:64])
    assert symbols is None or (isinstance(symbols, list) and len(symbols) == len(indexes) and all(type(v) is int for v in symbols[:64]))
    assert isinstance(cdf, list) and isinstance(cdf[0], list) and type(cdf[0][0]) is int
    assert len(cdf) == len(cdf_lengths) == len(offsets) and all(type(v) is int for v in cdf_lengths + offsets)
    assert 0 <= min(indexes) and max(indexes) < len(cdf)
    assert all(cdf[i][cdf_lengths[i] - 1] == 1 << 16 for i in (0, len(cdf) - 1))


class _Enc:
    def encode_with_indexes(self, symbols, indexes, cdf, cdf_lengths, offsets):
        _check_args(symbols, indexes, cdf, cdf_lengths, offsets)
        return pickle.dumps((symbols, indexes))


class _BufEnc:
    def __init__(self):
        self.sym, self.idx = [], []

    def encode_with_indexes(self, symbols, indexes, cdf, cdf_lengths, offsets):
        _check_args(symbols, indexes, cdf, cdf_lengths, offsets)
        self.sym += symbols
        self.idx += indexes

    def flush(self):
        return pickle.dumps((self.sym, self.idx))


class _Dec:
    def decode_with_indexes(self, string, indexes, cdf, cdf_lengths, offsets):
        _check_args(None, indexes, cdf, cdf_lengths, offsets)
        sym, idx = pickle.loads(string)
        assert idx == indexes, "decoder asked for other CDF indexes than the encoder used"
        return sym

    def set_stream(self, string):
        self.sym, self.idx = pickle.loads(string)
        self.pos = 0

    def decode_stream(self, indexes, cdf, cdf_lengths, offsets):
        _check_args(None, indexes, cdf, cdf_lengths, offsets)
        n = len(indexes)
        assert self.idx[self.pos:self.pos + n] == indexes, "decoder asked for other CDF indexes than the encoder used"
        out = self.sym[self.pos:self.pos + n]
        self.pos += n
        return out


@pytest.fixture()
def fake_ans(monkeypatch):
    pkg = sys.modules.get("compressai") or types.ModuleType("compressai")
    ans = types.ModuleType("compressai.ans")
    ans.RansEncoder, ans.RansDecoder, ans.BufferedRansEncoder = _Enc, _Dec, _BufEnc
    monkeypatch.setitem(sys.modules, "compressai", pkg)
    monkeypatch.setitem(sys.modules, "compressai.ans", ans)
    monkeypatch.setattr(pkg, "ans", ans, raising=False)
    return ans


def test_bitstream_entry_points_fail_loudly_without_tables_or_coder(synthetic_sd):
    m = resdsic_b200.WACNN().eval()
    m.load_state_dict(synthetic_sd, strict=True)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        m.compress(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        m.decompress([[b""], [b""]], (1, 1))
    if "compressai.ans" not in sys.modules:
        with pytest.raises(RuntimeError, match="compressai.ans"):
            m.entropy_bottleneck.entropy_coder.encode_with_indexes([], [], [], [], [])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_compress_decompress_round_trip(fake_ans, synthetic_sd, precision):
    model = resdsic_b200.WACNN.from_state_dict(synthetic_sd).to("cuda:0").eval().set_precision(precision)
    assert model.update() is True
    x = weights.make_image(2, 128, 64, seed=13).cuda()
    out = model.compress(x)
    assert list(out) == ["strings", "shape"] and tuple(out["shape"]) == (2, 1)
    y_strings, z_strings = out["strings"]
    assert len(y_strings) == 1 and len(z_strings) == 2 and all(isinstance(s, bytes) for s in y_strings + z_strings)
    sym, idx = pickle.loads(y_strings[0])
    assert len(sym) == len(idx) == 2 * 320 * 8 * 4
    want = model.symbols_and_indexes(x)
    # slice-major, then NCHW inside the slice (cnn.py:257-258)
    assert sym[:2 * 32 * 8 * 4] == want["y_symbols"][:, :32].reshape(-1).tolist()
    assert idx[-2 * 32 * 8 * 4:] == want["y_indexes"][:, 288:].reshape(-1).tolist()
    x_fwd = want["x_hat"].clamp(0, 1).clone()
    rec = model.decompress(out["strings"], out["shape"])
    assert list(rec) == ["x_hat"] and torch.equal(rec["x_hat"], x_fwd)
    # module-level API of the entropy models (entropy_models.py:205-287,512-526)
    eb, gc = model.entropy_bottleneck, model.gaussian_conditional
    z = weights.hash_symmetric("glue.z", (2, 192, 2, 3), 6.0).cuda()
    z_hat = eb.decompress(eb.compress(z), z.shape[-2:])
    med = eb._get_medians().detach().view(1, -1, 1, 1)
    assert torch.equal(z_hat, torch.round(z - med) + med)
    y, mu = weights.hash_symmetric("glue.y", (2, 32, 4, 4), 9.0).cuda(), weights.hash_symmetric("glue.mu", (2, 32, 4, 4), 2.0).cuda()
    sc = (weights.hash_uniform("glue.s", (2, 32, 4, 4)) * 20).cuda()
    ix = gc.build_indexes(sc)
    assert torch.equal(gc.decompress(gc.compress(y, ix, mu), ix, mu), torch.round(y - mu) + mu)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,quality", [("icd_gamma", 0.02), ("icd_gamma", 0.065), ("icd_gamma", 0.0035), ("imd_two", 0.065),
                                          ("cimd_gamma", 0.02), ("cimd_cat", 0.065), ("ind_md", 0.065), ("icd_nolrp", 0.065)])
def test_scalable_compress_decompress_round_trip(fake_ans, name, quality, precision):
    """ResDSIC scalable models (scalable/single_decoder.py:510-773): compress() hands the coder the symbols / indexes of
    `symbols_and_indexes`, in the reference's string layout; decompress() asks for exactly the encoder's CDF indexes
    (the stand-in decoder asserts it, slice by slice, both streams) and returns clamp(x_hat of the encoder pass)."""
    from tests.golden.make_golden_scalable import CASES, case_state_dict
    key, kw, _, (B, H, W) = CASES[name]
    m = resdsic_b200.models[key](N=192, M=320, **kw).eval()
    m.load_state_dict(case_state_dict(m.state_dict()), strict=True)
    m = m.to("cuda:0").set_precision(precision)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        m.compress(torch.zeros(B, 3, H, W, device="cuda:0"), quality=quality)
    m.update()
    x = weights.make_image(B, H, W, seed=5).cuda()
    q = m.lmbda_index_list[quality]
    out = m.compress(x, quality=quality)
    want = m.symbols_and_indexes(x, quality=quality)
    x_fwd = want["x_hat"].clamp(0, 1).clone()
    h, w = H // 16, W // 16
    assert list(out) == ["strings", "shape"] and len(out["strings"]) == (2 if q == 0 else 4) and len(out["shape"]) == (1 if q == 0 else 2)
    assert tuple(out["shape"][0]) == (h // 4, w // 4)
    sym, idx = pickle.loads(out["strings"][0][0])
    assert len(sym) == B * 320 * h * w and sym[:B * 32 * h * w] == want["y_symbols"][:, :32].reshape(-1).tolist()
    assert idx[-B * 32 * h * w:] == want["y_indexes"][:, 288:].reshape(-1).tolist()
    assert len(out["strings"][1]) == B
    if q != 0:
        assert len(out["strings"][2]) == B and len(out["strings"][3]) == 10 and all(len(s) == B for s in out["strings"][3])
        ps, pi = pickle.loads(out["strings"][3][7][B - 1])  # slice 7, last image
        assert ps == want["y_prog_symbols"][B - 1, 224:256].reshape(-1).tolist()
        assert pi == want["y_prog_indexes"][B - 1, 224:256].reshape(-1).tolist()
    rec = m.decompress(out["strings"], out["shape"], quality)
    assert list(rec) == ["x_hat"] and torch.equal(rec["x_hat"], x_fwd)
    # the quality may also be given as the level index (:658-661)
    assert torch.equal(m.decompress(out["strings"], out["shape"], q)["x_hat"], x_fwd)
