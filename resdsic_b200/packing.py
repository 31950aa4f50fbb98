"""One-time weight packing (host-side preprocessing, not on the per-image hot path).

Turns reference-layout parameters (OIHW conv weights, [Cin,Cout,k,k]
transposed-conv weights, [out,in] linears, GDN beta/gamma, EntropyBottleneck
matrices) into the kernel layouts of include/resdsic_b200.h.  Everything is
recomputed whenever the owning module's parameters change version or device.
"""
import math

import torch
import torch.nn.functional as F

from ._lib import EB_STRIDE


def _pad_rows(w2d):
    """The tensor-core kernel tiles N in multiples of 16: zero-pad bf16 weights to ceil16(Cout) rows."""
    if w2d.dtype == torch.bfloat16 and w2d.shape[0] % 16:
        w2d = F.pad(w2d, (0, 0, 0, 16 - w2d.shape[0] % 16))
    return w2d.contiguous()


def pack_conv_weight(w, dtype=torch.float32, k_pad_to=0):
    """[Cout,Cin,KH,KW] (nn.Conv2d) -> [Cout, KH*KW*Cin], tap-major / channel-minor."""
    Cout = w.shape[0]
    w2 = w.detach().permute(0, 2, 3, 1).reshape(Cout, -1)
    if k_pad_to and w2.shape[1] % k_pad_to:
        w2 = F.pad(w2, (0, k_pad_to - w2.shape[1] % k_pad_to))
    return _pad_rows(w2.to(dtype))


def pack_linear_weight(w, dtype=torch.float32):
    """nn.Linear [out,in] is already [N][K]."""
    return _pad_rows(w.detach().to(dtype))


# transposed conv k5 s2 p2 op1 (reference WACNN/utils.py:126-134): out[2j+py] gathers
# in[j + 1 - t] * w[.., kh = 2t + py]; as a stride-1 correlation over taps r (ascending
# input row j - pad + r): even phase (py=0): 3 taps, pad 1, kh = 4 - 2r;
# odd phase (py=1): 2 taps, pad 0, kh = 3 - 2r.   
DECONV_PHASES = {0: (3, 1, (4, 2, 0)), 1: (2, 0, (3, 1))}  # parity -> (taps, pad, kernel index per tap)


def pack_deconv_weight(w, dtype=torch.float32):
    """[Cin,Cout,5,5] -> dict[(py,px)] = (packed [Cout, R*S*Cin], R, S, pad_h, pad_w)."""
    assert w.shape[2:] == (5, 5)
    out = {}
    wd = w.detach()
    for py, (R, ph, khs) in DECONV_PHASES.items():
        for px, (S, pw, kws) in DECONV_PHASES.items():
            sub = wd[:, :, list(khs)][:, :, :, list(kws)]  # [Cin,Cout,R,S]
            packed = _pad_rows(sub.permute(1, 2, 3, 0).reshape(w.shape[1], -1).to(dtype))
            out[(py, px)] = (packed, R, S, ph, pw)
    return out


def pack_deconv_merged(w, dtype=torch.float32):
    """All four sub-pixel phases of the k5 s2 transposed conv as ONE 3x3 stride-1 pad-1 convolution with
    4*Cout output columns ordered n = c*4 + py*2 + px (so the PixelShuffle(2) store addressing scatters
    them): taps a phase does not use are zero.  Used for the 3-channel image head, where 4*Cout = 12 fits
    a single 16-wide MMA tile and the input is then read once instead of four times."""
    assert w.shape[2:] == (5, 5)
    Cin, Cout = w.shape[:2]
    wd = w.detach().float()
    out = torch.zeros(Cout, 2, 2, 3, 3, Cin, dtype=torch.float32, device=w.device)
    # row tap r of the 3x3 (input row j-1+r): even phase kh = 4-2r (r=0..2); odd phase kh = 5-2r (r=1,2)
    taps = {0: {0: 4, 1: 2, 2: 0}, 1: {1: 3, 2: 1}}
    for py in (0, 1):
        for px in (0, 1):
            for r, kh in taps[py].items():
                for s, kw in taps[px].items():
                    out[:, py, px, r, s, :] = wd[:, :, kh, kw].t()
    return _pad_rows(out.reshape(Cout * 4, 9 * Cin).to(dtype))


def nonneg_reparam(p, minimum):
    """NonNegativeParametrizer.forward (reference ops/parametrizers.py:46-49)."""
    pedestal = float(2 ** -18) ** 2
    bound = (minimum + pedestal) ** 0.5
    p = p.detach().float()
    return torch.clamp(p, min=torch.tensor(bound, dtype=torch.float32, device=p.device)) ** 2 - pedestal


def pack_gdn(beta, gamma, dtype=torch.float32):
    """GDN (reference layers/gdn.py:62-69): gamma' as a [C][C] 1x1 weight, beta' as its bias (fp32)."""
    return _pad_rows(nonneg_reparam(gamma, 0.0).to(dtype)), nonneg_reparam(beta, 1e-6).contiguous()


def pack_entropy_bottleneck(eb_params, quantiles):
    """[C][EB_STRIDE] fp32: softplus(matrix0..4) | bias0..4 | tanh(factor0..3) | median
    (reference entropy_models.py:401-420, :352-354).  softplus/tanh of the
    parameters are evaluated with the same torch ops the reference uses."""
    m = [F.softplus(eb_params[f"_matrix{i}"].detach().float()) for i in range(5)]
    b = [eb_params[f"_bias{i}"].detach().float() for i in range(5)]
    f = [torch.tanh(eb_params[f"_factor{i}"].detach().float()) for i in range(4)]
    Cn = m[0].shape[0]
    assert m[0].shape[1:] == (3, 1) and m[1].shape[1:] == (3, 3) and m[4].shape[1:] == (1, 3), \
        "kernel supports the reference's filters=(3,3,3,3)"
    cols = [t.reshape(Cn, -1) for t in (*m, *b, *f)]
    cols.append(quantiles.detach().float()[:, 0, 1:2])
    packed = torch.cat(cols, dim=1)
    assert packed.shape[1] == 59
    return F.pad(packed, (0, EB_STRIDE - packed.shape[1])).contiguous()


def scale_table(lo=0.11, hi=256.0, levels=64):
    """get_scale_table (reference cnn.py:14-20); evaluated on the CPU so the 64
    thresholds are bit-identical to the reference's on any device."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))
