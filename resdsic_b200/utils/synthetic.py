"""Deterministic, platform-independent synthetic weights and images (no compute of the
hot path lives here): the benchmark's synthetic workload and the parity tests' inputs.
`oracle/weights.py` re-exports this module for the test infrastructure.

The 75 M-parameter state_dict cannot be committed as a fixture, and libm /
torch RNG streams are not guaranteed bit-stable across hosts, so every tensor
is generated from an *integer* counter hash (splitmix64) of (tensor name,
element index): pure uint64 arithmetic in numpy, then an exact int->float32
conversion.  The same call therefore yields bit-identical tensors in the
build container (where the reference produces the golden outputs) and on the
GPU box (where the oracle and the CUDA path consume them).

`state_dict_spec()` restates the reference's parameter/buffer inventory
(names + shapes; SURVEY.md section 8b, reference `cnn.py:31-132`,
`layers/layers.py:45-81`, `layers/win_attention.py:58-79`, `layers/gdn.py:52-60`,
`entropy_models/entropy_models.py:98-100,325-350,563-572`).  make_golden.py
asserts it equals the reference's own `state_dict()` key-for-key.
"""
import math
import re
import zlib
from collections import OrderedDict

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_uniform(name: str, shape, seed: int = 0) -> torch.Tensor:
    """U[0,1) float32 tensor, exactly reproducible (24-bit mantissa draws)."""
    n = int(np.prod(shape)) if len(shape) else 1
    key = np.uint64(zlib.crc32(name.encode()) | (int(seed) << 32))
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + _splitmix64(np.array([key], dtype=np.uint64))[0]
    bits = _splitmix64(ctr) >> np.uint64(40)  # top 24 bits
    u = bits.astype(np.float32) * np.float32(1.0 / (1 << 24))  # exact
    return torch.from_numpy(u.reshape(tuple(shape)))


def hash_symmetric(name, shape, amp: float, seed: int = 0) -> torch.Tensor:
    """U(-amp, amp): (2u-1)*amp, all fp32 ops exact/deterministic."""
    return (hash_uniform(name, shape, seed) * 2.0 - 1.0) * float(amp)


# --------------------------------------------------------------------------
# state_dict inventory
# --------------------------------------------------------------------------
def _ru(prefix, C, spec):
    h = C // 2
    spec[f"{prefix}.conv.0.weight"] = (h, C, 1, 1)
    spec[f"{prefix}.conv.0.bias"] = (h,)
    spec[f"{prefix}.conv.2.weight"] = (h, h, 3, 3)
    spec[f"{prefix}.conv.2.bias"] = (h,)
    spec[f"{prefix}.conv.4.weight"] = (C, h, 1, 1)
    spec[f"{prefix}.conv.4.bias"] = (C,)


def _attn_block(prefix, C, ws, heads, spec):
    for i in range(3):
        _ru(f"{prefix}.conv_a.{i}", C, spec)
    a = f"{prefix}.conv_b.0.attn"
    spec[f"{a}.relative_position_bias_table"] = ((2 * ws - 1) ** 2, heads)
    spec[f"{a}.relative_position_index"] = (ws * ws, ws * ws)
    spec[f"{a}.qkv.weight"] = (3 * C, C)
    spec[f"{a}.qkv.bias"] = (3 * C,)
    spec[f"{a}.proj.weight"] = (C, C)
    spec[f"{a}.proj.bias"] = (C,)
    for i in (1, 2, 3):
        _ru(f"{prefix}.conv_b.{i}", C, spec)
    spec[f"{prefix}.conv_b.4.weight"] = (C, C, 1, 1)
    spec[f"{prefix}.conv_b.4.bias"] = (C,)


def _gdn(prefix, C, spec):
    spec[f"{prefix}.beta"] = (C,)
    spec[f"{prefix}.gamma"] = (C, C)
    spec[f"{prefix}.beta_reparam.pedestal"] = (1,)
    spec[f"{prefix}.beta_reparam.lower_bound.bound"] = (1,)
    spec[f"{prefix}.gamma_reparam.pedestal"] = (1,)
    spec[f"{prefix}.gamma_reparam.lower_bound.bound"] = (1,)


def _conv(prefix, cout, cin, k, spec):
    spec[f"{prefix}.weight"] = (cout, cin, k, k)
    spec[f"{prefix}.bias"] = (cout,)


def state_dict_spec(N=192, M=320):
    """OrderedDict name -> shape, in the reference's registration order."""
    s = OrderedDict()
    _conv("g_a.0", N, 3, 5, s)
    _gdn("g_a.1", N, s)
    _conv("g_a.2", N, N, 5, s)
    _gdn("g_a.3", N, s)
    _attn_block("g_a.4", N, 8, 8, s)
    _conv("g_a.5", N, N, 5, s)
    _gdn("g_a.6", N, s)
    _conv("g_a.7", M, N, 5, s)
    _attn_block("g_a.8", M, 4, 8, s)
    _attn_block("g_s.0", M, 4, 8, s)
    # ConvTranspose2d weight is [Cin, Cout, k, k]  (WACNN/utils.py:126-134)
    s["g_s.1.weight"], s["g_s.1.bias"] = (M, N, 5, 5), (N,)
    _gdn("g_s.2", N, s)
    s["g_s.3.weight"], s["g_s.3.bias"] = (N, N, 5, 5), (N,)
    _gdn("g_s.4", N, s)
    _attn_block("g_s.5", N, 8, 8, s)
    s["g_s.6.weight"], s["g_s.6.bias"] = (N, N, 5, 5), (N,)
    _gdn("g_s.7", N, s)
    s["g_s.8.weight"], s["g_s.8.bias"] = (N, 3, 5, 5), (3,)
    for i, (co, ci) in zip((0, 2, 4, 6, 8), ((320, 320), (288, 320), (256, 288), (224, 256), (192, 224))):
        _conv(f"h_a.{i}", co, ci, 3, s)
    for h in ("h_mean_s", "h_scale_s"):
        _conv(f"{h}.0", 192, 192, 3, s)
        _conv(f"{h}.2.0", 224 * 4, 192, 3, s)
        _conv(f"{h}.4", 256, 224, 3, s)
        _conv(f"{h}.6.0", 288 * 4, 256, 3, s)
        _conv(f"{h}.8", 320, 288, 3, s)
    chain = (224, 176, 128, 64, 32)
    for fam, extra in (("cc_mean_transforms", 0), ("cc_scale_transforms", 0), ("lrp_transforms", 1)):
        for i in range(10):
            cin = 320 + 32 * min(i + extra, 5 + extra)
            for j, co in zip((0, 2, 4, 6, 8), chain):
                _conv(f"{fam}.{i}.{j}", co, cin, 3, s)
                cin = co
    eb = "entropy_bottleneck"
    filt = (1, 3, 3, 3, 3, 1)
    for i in range(5):
        s[f"{eb}._matrix{i}"] = (N, filt[i + 1], filt[i])
        s[f"{eb}._bias{i}"] = (N, filt[i + 1], 1)
        if i < 4:
            s[f"{eb}._factor{i}"] = (N, filt[i + 1], 1)
    s[f"{eb}.quantiles"] = (N, 1, 3)
    for b in ("_offset", "_quantized_cdf", "_cdf_length"):
        s[f"{eb}.{b}"] = (0,)
    s[f"{eb}.target"] = (3,)
    s[f"{eb}.likelihood_lower_bound.bound"] = (1,)
    gc = "gaussian_conditional"
    for b in ("_offset", "_quantized_cdf", "_cdf_length", "scale_table"):
        s[f"{gc}.{b}"] = (0,)
    s[f"{gc}.scale_bound"] = (1,)
    s[f"{gc}.likelihood_lower_bound.bound"] = (1,)
    s[f"{gc}.lower_bound_scale.bound"] = (1,)
    return s


def relative_position_index(ws: int) -> torch.Tensor:
    """index[i,j] = (hi-hj+ws-1)(2ws-1) + (wi-wj+ws-1)  (win_attention.py:64-74)."""
    t = torch.arange(ws * ws)
    hi, wi = t // ws, t % ws
    return (hi[:, None] - hi[None, :] + ws - 1) * (2 * ws - 1) + (wi[:, None] - wi[None, :] + ws - 1)


def scale_table(lo=0.11, hi=256.0, levels=64) -> torch.Tensor:
    """exp(linspace(ln lo, ln hi, levels)) -- cnn.py:14-20.  The golden fixture
    pins these 64 floats bit-for-bit."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


# per-layer amplitude multipliers that keep the synthetic latents lively:
# y spans several quantisation bins, z spans several integers, scales spread.
_GAINS = (
    (r"^g_[as]\.", 0.7071),  # analysis/synthesis: var 1/fan_in (GDN + residual units keep the level)
    (r"^g_a\.7\.weight$", 2.5),
    (r"^h_a\.8\.weight$", 3.0),
    (r"^h_(mean|scale)_s\.8\.weight$", 1.5),
    (r"^cc_scale_transforms\.\d\.8\.weight$", 1.5),
    (r"^cc_mean_transforms\.\d\.8\.weight$", 0.4),
    (r"^g_s\.[136]\.weight$", 0.25),  # IGDN grows quadratically: keep x_hat O(1)
    (r"^g_s\.8\.weight$", 0.35),
)


# "lowrate" profile: the same weights with the latent / hyper-latent amplitudes scaled down, so that most
# symbols are zero (bpp ~4 instead of ~14 on the synthetic images) -- the regime a trained codec operates in,
# and the one on which the bf16 tolerances of BASELINE.json (x_hat 1e-2, bpp 0.1 %, PSNR 0.02 dB) are asserted.
_LOWRATE = (
    (r"^g_a\.7\.weight$", 0.25),
    (r"^h_a\.8\.weight$", 0.5),
    (r"^h_(mean|scale)_s\.8\.weight$", 0.5),
    (r"^cc_(mean|scale)_transforms\.\d\.8\.weight$", 0.5),
)


def _layer_gain(name, profile="stress"):
    g = 1.0
    for pat, v in _GAINS:
        if re.match(pat, name):
            g = v  # last match wins
    if profile == "lowrate":
        for pat, v in _LOWRATE:
            if re.match(pat, name):
                g *= v
    return g


def make_state_dict(seed: int = 0, N=192, M=320, gain: float = 1.0, profile: str = "stress"):
    """Synthetic but *non-degenerate* weights: conv/linear weights are uniform
    with Kaiming-like variance (so activations neither vanish nor explode),
    biases are non-zero, GDN beta/gamma and the EntropyBottleneck parameters
    are the reference initial values plus a perturbation, so every term of the
    arithmetic is exercised."""
    spec = state_dict_spec(N, M)
    sd = OrderedDict()
    pedestal = float(2 ** -18) ** 2
    for name, shape in spec.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "relative_position_index":
            ws = int(round(math.sqrt(shape[0])))
            t = relative_position_index(ws)
        elif leaf == "relative_position_bias_table":
            t = hash_symmetric(name, shape, 0.5, seed)
        elif leaf == "pedestal":
            t = torch.tensor([pedestal], dtype=torch.float32)
        elif leaf == "bound" and "beta_reparam" in name:
            t = torch.tensor([(1e-6 + pedestal) ** 0.5], dtype=torch.float32)
        elif leaf == "bound" and "gamma_reparam" in name:
            t = torch.tensor([(0.0 + pedestal) ** 0.5], dtype=torch.float32)
        elif leaf == "beta":
            b = 1.0 + hash_symmetric(name, shape, 0.3, seed)
            t = torch.sqrt(torch.clamp(b + pedestal, min=pedestal))
        elif leaf == "gamma":
            g = 0.1 * torch.eye(shape[0]) + hash_uniform(name, shape, seed) * 0.004
            t = torch.sqrt(torch.clamp(g + pedestal, min=pedestal))
        elif name.startswith("entropy_bottleneck."):
            if leaf.startswith("_matrix"):
                i = int(leaf[-1])
                filt = (1, 3, 3, 3, 3, 1)
                scale = 10.0 ** (1 / 5)
                init = math.log(math.expm1(1 / scale / filt[i + 1]))
                t = init + hash_symmetric(name, shape, 0.2, seed)
            elif leaf.startswith("_bias"):
                t = hash_symmetric(name, shape, 0.5, seed)
            elif leaf.startswith("_factor"):
                t = hash_symmetric(name, shape, 0.3, seed)
            elif leaf == "quantiles":
                med = hash_symmetric(name, (shape[0], 1, 1), 0.4, seed)
                t = torch.cat([med - 10.0, med, med + 10.0], dim=2)
            elif leaf == "target":
                v = math.log(2 / 1e-9 - 1)
                t = torch.tensor([-v, 0.0, v], dtype=torch.float32)
            elif leaf == "bound":
                t = torch.tensor([1e-9], dtype=torch.float32)
            else:
                t = torch.zeros(0, dtype=torch.int32)
        elif name.startswith("gaussian_conditional."):
            if leaf == "scale_table":
                t = torch.zeros(0, dtype=torch.float32)
            elif leaf == "scale_bound" or name.endswith("lower_bound_scale.bound"):
                t = torch.tensor([0.11], dtype=torch.float32)
            elif leaf == "bound":
                t = torch.tensor([1e-9], dtype=torch.float32)
            else:
                t = torch.zeros(0, dtype=torch.int32)
        elif leaf == "bias" and name.startswith("cc_scale_transforms.") and name.endswith(".8.bias"):
            # spread the predicted scales over the 64-entry table (most below ~4, a tail above)
            t = hash_uniform(name, shape, seed) ** 3 * 12.0 - 0.3
        elif leaf == "bias":
            t = hash_symmetric(name, shape, 0.05, seed)
        elif leaf == "weight":
            if len(shape) == 4:
                # deconv weights are [Cin,Cout,k,k]: fan-in there is Cin*k*k/stride^2
                is_deconv = name.startswith("g_s.") and name.split(".")[1] in ("1", "3", "6", "8") and len(name.split(".")) == 3
                fan_in = (shape[0] * shape[2] * shape[3] / 4.0) if is_deconv else shape[1] * shape[2] * shape[3]
            else:
                fan_in = shape[1]
            amp = gain * _layer_gain(name, profile) * math.sqrt(6.0 / fan_in)  # var = 2/fan_in (GELU nets)
            t = hash_symmetric(name, shape, amp, seed)
        else:
            raise KeyError(name)
        assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
        sd[name] = t.contiguous()
    return sd


def make_image(B, H, W, seed: int = 0) -> torch.Tensor:
    """Synthetic image batch in [0,1): 8x8 blocky low-frequency content plus
    fine noise, so that latents/scales are spread rather than degenerate."""
    lo = hash_uniform(f"img.lo.{B}x{H}x{W}", (B, 3, (H + 7) // 8, (W + 7) // 8), seed)
    lo = lo.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :H, :W]
    hi = hash_uniform(f"img.hi.{B}x{H}x{W}", (B, 3, H, W), seed)
    return (lo * 0.75 + hi * 0.25).contiguous()


# --------------------------------------------------------------------------
# generic synthetic weights for builder-defined models (stf): keyed on the module's own state_dict
# --------------------------------------------------------------------------
_STF_GAINS = (
    (r"^g_a\.layers\.3\.blocks\.1\.mlp\.2\.weight$", 6.0),
    (r"^h_a\.8\.weight$", 3.0),
    (r"^h_(mean|scale)_s\.8\.weight$", 1.5),
    (r"^cc_scale_transforms\.\d+\.4\.weight$", 1.5),
    (r"^cc_mean_transforms\.\d+\.4\.weight$", 0.4),
    (r"^g_s\.end_conv\.", 0.5),
)


def synth_state_dict(reference_sd, seed: int = 0, gains=_STF_GAINS):
    """Hash-seeded values for every entry of `reference_sd` (name -> tensor giving shape / dtype), using the
    same conventions as make_state_dict: He-like uniform weights, small non-zero biases, LayerNorm weights
    around 1, perturbed EntropyBottleneck parameters.  Constant buffers are copied through."""
    sd = OrderedDict()
    for name, ref in reference_sd.items():
        shape, leaf = tuple(ref.shape), name.rsplit(".", 1)[-1]
        g = 1.0
        for pat, v in gains:
            if re.match(pat, name):
                g = v
        if ref.numel() == 0 or leaf in ("pedestal", "bound", "target", "relative_position_index", "scale_bound"):
            t = ref.clone()
        elif name.split(".")[0].startswith("entropy_bottleneck"):  # (also entropy_bottleneck_prog of the scalable models)
            if leaf.startswith("_matrix"):
                t = ref.clone() + hash_symmetric(name, shape, 0.2, seed)
            elif leaf.startswith("_bias"):
                t = hash_symmetric(name, shape, 0.5, seed)
            elif leaf.startswith("_factor"):
                t = hash_symmetric(name, shape, 0.3, seed)
            elif leaf == "quantiles":
                med = hash_symmetric(name, (shape[0], 1, 1), 0.4, seed)
                t = torch.cat([med - 10.0, med, med + 10.0], dim=2)
            else:
                t = ref.clone()
        elif leaf == "relative_position_params" or leaf == "relative_position_bias_table":
            t = hash_symmetric(name, shape, 0.5, seed)
        elif leaf == "weight" and len(shape) == 1:  # LayerNorm
            t = 1.0 + hash_symmetric(name, shape, 0.2, seed)
        elif leaf == "bias":
            t = hash_symmetric(name, shape, 0.1 if ".norm" in name or ".ln" in name else 0.05, seed)
        elif leaf == "weight":
            fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            var2 = len(shape) == 4  # conv stacks are GELU nets: var 2/fan_in; linears: var 1/fan_in
            t = hash_symmetric(name, shape, g * math.sqrt((6.0 if var2 else 3.0) / fan_in), seed)
        else:
            raise KeyError(name)
        assert tuple(t.shape) == shape, (name, t.shape, shape)
        sd[name] = t.to(ref.dtype).contiguous()
    return sd


# --------------------------------------------------------------------------
# "refinit": the reference constructor's own random init (north_star: parity on random-init weights)
# --------------------------------------------------------------------------
def refinit_model(seed: int = 0, N=192, M=320):
    """`resdsic_b200.WACNN(N, M)` constructed under `torch.manual_seed(seed)` on the CPU generator.  The
    constructor makes the same init calls in the same order as the reference's (`cnn.py:26-132`; its Kaiming
    loop is a no-op, see layers/conv.py:_default_conv_init), so this draws the SAME 75 M weights as
    `torch.manual_seed(seed); compress.models.WACNN(N, M)` -- asserted tensor for tensor against the imported
    reference by tests/golden/make_golden_refinit.py, and pinned on any host by the checksums that script
    commits (tests/test_refinit.py).  The caller's global RNG state is preserved."""
    from ..models import WACNN
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        return WACNN(N=N, M=M)


def refinit_state_dict(seed: int = 0, N=192, M=320):
    return OrderedDict((k, v.detach().clone()) for k, v in refinit_model(seed, N, M).state_dict().items())


def rand_image(B, H, W, seed: int = 1) -> torch.Tensor:
    """`torch.rand(B,3,H,W)` (the synthetic input of the BASELINE configs) from a private CPU
    generator, i.e. independent of the global RNG state."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand(B, 3, H, W, generator=g)


def tensor_checksum(t) -> np.ndarray:
    """Three exact-ish numbers pinning a tensor's contents (float64 sum, strided abs-sum, last element)."""
    t = t.detach().double().reshape(-1)
    if t.numel() == 0:
        return np.zeros(3)
    return np.array([t.sum().item(), t[:: max(1, t.numel() // 7)][:7].abs().sum().item(), float(t[-1])])
