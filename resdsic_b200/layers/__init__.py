from .base import GELU, B200Module, Ctx, Sequential
from .conv import (Conv2d, ConvTranspose2d, Linear, PixelShuffle, SubpelConv, conv, conv1x1, conv3x3, deconv,
                   subpel_conv3x3)
from .gdn import GDN
from .layers import ResidualUnit, Win_noShift_Attention
from .swin import Block as SwinBlock
from .swin import WMSA, BasicLayer, LayerNorm, PatchMerging, PatchSplit
from .win_attention import WinBasedAttention, WindowAttention

__all__ = ["GELU", "B200Module", "Ctx", "Sequential", "Conv2d", "ConvTranspose2d", "Linear", "PixelShuffle",
           "SubpelConv", "conv", "conv1x1", "conv3x3", "deconv", "subpel_conv3x3", "GDN", "ResidualUnit",
           "Win_noShift_Attention", "WinBasedAttention", "WindowAttention", "SwinBlock", "LayerNorm", "WMSA", "BasicLayer", "PatchMerging", "PatchSplit"]
