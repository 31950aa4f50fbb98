"""Host-side helpers mirroring the reference callers' conventions."""
import torch
import torch.nn.functional as F

from .pipeline import ForwardPipeline  # noqa: F401
from .sharding import gather_to_rank0, max_over_ranks, shard_range  # noqa: F401


def pad_to_multiple(x, m=64):
    """Caller padding rule (reference utils/eval_model/__main__.py:89-101, training/step.py:236-238):
    symmetric zero padding of H, W up to a multiple of 64.  Returns (padded, (left, right, top, bottom))."""
    H, W = x.shape[-2:]
    Hp, Wp = -(-H // m) * m, -(-W // m) * m
    left, top = (Wp - W) // 2, (Hp - H) // 2
    pad = (left, Wp - W - left, top, Hp - H - top)
    return F.pad(x, pad, mode="constant", value=0), pad


def crop(x, pad):
    left, right, top, bottom = pad
    return F.pad(x, (-left, -right, -top, -bottom))


def bpp(likelihoods, num_pixels):
    """reference training/loss.py:14-22."""
    import math
    return sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in likelihoods.values())


def psnr(a, b):
    """reference utils/functions.py:55-57."""
    return -10.0 * torch.log10(torch.mean((a - b) ** 2))
