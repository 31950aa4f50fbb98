"""STF (`-m stf`): a symmetrical Swin-transformer codec on the B200 kernel library.

The reference tree ships NO STF model (no `stf.py`, `-m stf` is not a valid choice there), so
this architecture is **builder-defined** and its model-level parity is **unpinned**: it follows the upstream
STF design the reference's README describes as far as it can be reconstructed (patch size 2,
embed dim 48, depths [2,2,6,2], heads [3,6,12,24], window 4, 384 latent channels, 12 channel slices with 6
support slices, 3-conv context transforms).  What IS pinned is every building block: the Swin block is the
reference's `TCM/tcm.py:Block` (tests pin it against the reference module), window attention / entropy
models / slice-context loop are the reference's `cnn` ones.  The model-level oracle is
`oracle/stf_oracle.py` (a plain PyTorch composition of those pinned blocks).

It reuses `WACNN`'s whole forward machinery (program builder, slice loop, CUDA-graph execution): only the
transforms differ.
"""
import torch
import torch.nn as nn

from ..entropy_models import EntropyBottleneck, GaussianConditional
from ..layers import GELU, B200Module, BasicLayer, Conv2d, LayerNorm, Sequential, conv, conv3x3, subpel_conv3x3
from ..layers.conv import SubpelConv
from .wacnn import WACNN, CompressionModel


class _Analysis(B200Module):
    """patch embed (conv k2 s2 + LN) -> 4 Swin stages with 3 patch mergings: [B,3,H,W] -> [B,H/16,W/16,8C]."""

    def __init__(self, embed_dim, depths, heads, window):
        super().__init__()
        self.proj = Conv2d(3, embed_dim, 2, 2, padding=0)
        self.norm = LayerNorm(embed_dim)
        self.layers = nn.ModuleList(
            BasicLayer(embed_dim * 2 ** i, depths[i], heads[i], window, "down" if i < len(depths) - 1 else None)
            for i in range(len(depths)))

    def emit(self, ctx, x, last_kw=None, **kw):
        t = self.norm.emit(ctx, self.proj.emit(ctx, x))
        layers = list(self.layers)
        for i, layer in enumerate(layers):
            t = layer.emit(ctx, t, **(dict(last_kw or {}) if i + 1 == len(layers) else {}))
        return t


class _Synthesis(B200Module):
    """mirror: 4 Swin stages with 3 patch splits, then conv5x5 -> PixelShuffle(2) -> conv3x3 to RGB."""

    def __init__(self, embed_dim, depths, heads, window):
        super().__init__()
        n = len(depths)
        self.layers = nn.ModuleList(
            BasicLayer(embed_dim * 2 ** (n - 1 - i), depths[n - 1 - i], heads[n - 1 - i], window, "up" if i < n - 1 else None)
            for i in range(n))
        self.end_conv = Sequential(SubpelConv(embed_dim, embed_dim, 2, kernel_size=5), Conv2d(embed_dim, 3, 3))

    def emit(self, ctx, x, last_kw=None, **kw):
        t = x
        for layer in self.layers:
            t = layer.emit(ctx, t)
        return self.end_conv.emit(ctx, t, last_kw=last_kw)


def _cc3(cin):
    return Sequential(conv(cin, 224, stride=1, kernel_size=3), GELU(), conv(224, 128, stride=1, kernel_size=3), GELU(),
                      conv(128, 32, stride=1, kernel_size=3))


class SymmetricalTransFormer(WACNN):
    """STF.  forward(x) -> {"x_hat", "likelihoods": {"y", "z"}} like every registry model."""

    train_forward_impl = None  # no differentiable training forward for this model: train() gives forward values only

    def __init__(self, N=192, M=384, embed_dim=48, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=4,
                 num_slices=12, max_support_slices=6, **kwargs):
        CompressionModel.__init__(self)
        if M != embed_dim * 2 ** (len(depths) - 1):
            raise ValueError("M must equal embed_dim * 2**(stages-1)")
        if M % num_slices or (M // num_slices) % 16:
            raise ValueError("slice width must be a multiple of 16 channels")
        self.N, self.M = N, M
        self.num_slices, self.max_support_slices = num_slices, max_support_slices
        self.slice_channels = sc = M // num_slices
        self.g_a = _Analysis(embed_dim, depths, num_heads, window_size)
        self.g_s = _Synthesis(embed_dim, depths, num_heads, window_size)
        self.h_a = Sequential(conv3x3(M, 384), GELU(), conv3x3(384, 336), GELU(), conv3x3(336, 288, stride=2), GELU(),
                              conv3x3(288, 240), GELU(), conv3x3(240, N, stride=2))

        def h_s():
            return Sequential(conv3x3(N, 240), GELU(), subpel_conv3x3(240, 288, 2), GELU(), conv3x3(288, 336), GELU(),
                              subpel_conv3x3(336, 384, 2), GELU(), conv3x3(384, M))

        self.h_mean_s = h_s()
        self.h_scale_s = h_s()
        S = max_support_slices
        self.cc_mean_transforms = nn.ModuleList(_cc3(M + sc * min(i, S)) for i in range(num_slices))
        self.cc_scale_transforms = nn.ModuleList(_cc3(M + sc * min(i, S)) for i in range(num_slices))
        self.lrp_transforms = nn.ModuleList(_cc3(M + sc * min(i + 1, S + 1)) for i in range(num_slices))
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)
        self._init_runtime()  # plans, noise_override, decoder plans ... (everything WACNN's forward / train / decode paths use)

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls()
        net.load_state_dict(state_dict)
        return net
