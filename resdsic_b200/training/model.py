"""Differentiable training forward of WACNN (`-m cnn`): the reference's `WACNN.forward` in train() mode
(models/WACNN/cnn.py:143-193) composed from the autograd nodes of functions.py over the SAME module parameters
(`resdsic_b200.WACNN`), so `train.py`'s optimiser / clipping / aux-optimiser code runs unchanged on the result.

Parameter-sized preprocessing (GDN reparametrisation with the LowerBound rule, softplus / tanh of the
EntropyBottleneck matrices / factors) is done with torch ops exactly as the reference does it (gdn.py:62-69,
entropy_models.py:401-420): its autograd carries the kernels' packed-parameter gradients back to the parameters.
"""
import torch
import torch.nn.functional as F

from ..layers import GDN, GELU, Win_noShift_Attention
from ..layers.conv import Conv2d, ConvTranspose2d, SubpelConv
from .._lib import EB_STRIDE
from ..ops import LowerBoundFunction
from . import functions as Fn


def _conv(m, x):
    return Fn.ConvFn.apply(x, m.weight, m.bias, m.stride, m.padding)


def _linear(m, x):
    return Fn.ConvFn.apply(x, m.weight[:, :, None, None], m.bias, 1, 0)


def _reparam(rp, p):
    """NonNegativeParametrizer.forward (ops/parametrizers.py:46-49) with the LowerBound gradient rule."""
    return LowerBoundFunction.apply(p, rp.lower_bound.bound) ** 2 - rp.pedestal


def _gdn(m, x):
    """layers/gdn.py:62-75."""
    beta = _reparam(m.beta_reparam, m.beta)
    gamma = _reparam(m.gamma_reparam, m.gamma)
    norm = Fn.ConvFn.apply(Fn.SquareFn.apply(x), gamma[:, :, None, None], beta, 1, 0)
    return Fn.GdnScaleFn.apply(x, norm, m.inverse)


def _residual_unit(m, x):
    """layers/layers.py:52-71."""
    t = Fn.GeluFn.apply(_conv(m.conv[0], x))
    t = Fn.GeluFn.apply(_conv(m.conv[2], t))
    t = _conv(m.conv[4], t)
    return Fn.GeluFn.apply(Fn.AddFn.apply(t, x))


def _win_based_attention(m, x):
    """layers/win_attention.py:153-207 (+ WindowAttention.forward :84-115)."""
    a = m.attn
    qkv = _linear(a.qkv, x)
    att = Fn.WindowAttentionFn.apply(qkv, a.relative_position_bias_table, a.num_heads, a.window_size[0], m.shift_size, float(a.scale))
    return Fn.AddFn.apply(x, _linear(a.proj, att))


def _attention_block(m, x):
    """Win_noShift_Attention.forward (layers/layers.py:83-89)."""
    a = x
    for ru in m.conv_a:
        a = _residual_unit(ru, a)
    b = _win_based_attention(m.conv_b[0], x)
    for ru in list(m.conv_b)[1:4]:
        b = _residual_unit(ru, b)
    b = _conv(m.conv_b[4], b)
    return Fn.GateFn.apply(a, b, x)


def _apply(m, x):
    if isinstance(m, Conv2d):
        return _conv(m, x)
    if isinstance(m, ConvTranspose2d):
        return Fn.DeconvFn.apply(x, m.weight, m.bias)
    if isinstance(m, SubpelConv):
        return Fn.PixelShuffleFn.apply(_conv(m[0], x))
    if isinstance(m, GDN):
        return _gdn(m, x)
    if isinstance(m, GELU):
        return Fn.GeluFn.apply(x)
    if isinstance(m, Win_noShift_Attention):
        return _attention_block(m, x)
    raise TypeError(f"no training node for {type(m).__name__}")


def _seq(seq, x):
    for m in seq:
        x = _apply(m, x)
    return x


def packed_entropy_bottleneck(eb):
    """[C][EB_STRIDE]: softplus(matrix0..4) | bias0..4 | tanh(factor0..3) | median, WITH autograd history."""
    m = [F.softplus(getattr(eb, f"_matrix{i}")) for i in range(5)]
    b = [getattr(eb, f"_bias{i}") for i in range(5)]
    f = [torch.tanh(getattr(eb, f"_factor{i}")) for i in range(4)]
    Cn = m[0].shape[0]
    cols = [t.reshape(Cn, -1) for t in (*m, *b, *f)] + [eb.quantiles[:, 0, 1:2].detach()]
    packed = torch.cat(cols, dim=1)
    return F.pad(packed, (0, EB_STRIDE - packed.shape[1]))


def train_forward(model, x, noise=None):
    """x: [B,3,H,W] in [0,1] on a CUDA device, H and W multiples of 64.  `noise` = {"y": [B,320,h,w], "z":
    [B,192,h/4,w/4]} injects the U(-1/2,1/2) draws (tests); otherwise they are drawn on the device.
    Returns the reference's output dict with autograd history."""
    if not x.is_cuda:
        raise RuntimeError("resdsic_b200 training runs on CUDA devices only (no CPU fallback)")
    if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 64 or x.shape[3] % 64:
        raise ValueError(f"expected [B,3,H,W] with H, W multiples of 64, got {tuple(x.shape)}")
    gc, eb = model.gaussian_conditional, model.entropy_bottleneck
    xl = Fn.ToChannelsLastFn.apply(x.float())
    y = _seq(model.g_a, xl)                                   # [B,h,w,320]
    z = _seq(model.h_a, y)                                    # [B,h/4,w/4,192]
    B, h, w, M = y.shape

    def draw(key, ref):
        if noise is not None:
            return Fn.ToChannelsLastFn.apply(noise[key].to(x.device, torch.float32))
        return torch.empty_like(ref).uniform_(-0.5, 0.5)

    z_lik, z_hat = Fn.EntropyBottleneckFn.apply(z, packed_entropy_bottleneck(eb), draw("z", z), float(eb.likelihood_bound))
    latent_scales = _seq(model.h_scale_s, z_hat)
    latent_means = _seq(model.h_mean_s, z_hat)

    ny = draw("y", y)
    from ..models.wacnn import get_scale_table  # (the table only feeds build_indexes, which training does not use)
    table = gc.scale_table if gc.scale_table.numel() >= 2 else get_scale_table()
    table = table.detach().float().contiguous().to(x.device)
    c = model.slice_channels
    y_hat_slices, y_liks = [], []
    for i in range(model.num_slices):
        support = y_hat_slices if model.max_support_slices < 0 else y_hat_slices[:model.max_support_slices]
        mean_support = Fn.CatChannelsFn.apply(latent_means, *support) if support else latent_means
        scale_support = Fn.CatChannelsFn.apply(latent_scales, *support) if support else latent_scales
        mu = _seq(model.cc_mean_transforms[i], mean_support)
        scale = _seq(model.cc_scale_transforms[i], scale_support)
        y_slice = Fn.SliceChannelsFn.apply(y, c * i, c)
        n_slice = Fn.SliceChannelsFn.apply(ny, c * i, c)
        lik, y_hat_slice = Fn.GaussianConditionalFn.apply(y_slice, mu, scale, n_slice, table, float(gc.scale_bound_value),
                                                          float(gc.likelihood_bound))
        y_liks.append(lik)
        lrp = _seq(model.lrp_transforms[i], Fn.CatChannelsFn.apply(mean_support, y_hat_slice))
        y_hat_slices.append(Fn.LrpFn.apply(y_hat_slice, lrp))
    y_hat = Fn.CatChannelsFn.apply(*y_hat_slices)
    x_hat = Fn.ToChannelsFirstFn.apply(_seq(model.g_s, y_hat))
    return {"x_hat": x_hat, "likelihoods": {"y": Fn.CatNchwFn.apply(*y_liks), "z": z_lik}}


def aux_loss_of(eb):
    """EntropyBottleneck.loss (entropy_models.py:396-399) with a gradient to `.quantiles` only."""
    packed = packed_entropy_bottleneck(eb).detach()
    return Fn.AuxLossFn.apply(eb.quantiles, packed, eb.target.float())


def aux_loss(model):
    """CompressionModel.aux_loss (WACNN/base.py:22-27): gradient to the `.quantiles` parameters (train.py:59-68)."""
    from ..entropy_models import EntropyBottleneck
    return sum(aux_loss_of(m) for m in model.modules() if isinstance(m, EntropyBottleneck))
