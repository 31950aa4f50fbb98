"""TEST INFRASTRUCTURE -- PyTorch (CPU, fp32) oracle of the builder-defined `stf` model
(resdsic_b200/models/stf.py).

PARITY UNPINNED at the model level: the reference tree contains no STF implementation, so
there is nothing of the reference to run for this architecture.  Every block used here IS pinned: the Swin
block is `wacnn_oracle.swin_block` (pinned to the reference's TCM/tcm.py:Block), the hyperprior entropy
models and the channel-slice context loop are the pinned `cnn` ones (cnn.py:143-193), generalised to
M = 384 / 12 slices / 6 support slices / 3-conv transforms as in upstream STF.
"""
import torch
import torch.nn.functional as F

from . import wacnn_oracle as O


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"])


def _stage(x, sd, p, depth, head_dim, ws):
    for j in range(depth):
        x = O.swin_block(x, sd, f"{p}.blocks.{j}", head_dim, ws, shifted=(j % 2 == 1))
    return x


def patch_merging(x, sd, p):
    """2x2 gather in (dy,dx) = (0,0),(0,1),(1,0),(1,1) order -> LN(4C) -> Linear(4C->2C, no bias)."""
    g = torch.cat([x[:, 0::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 0::2], x[:, 1::2, 1::2]], -1)
    return F.linear(_ln(g, sd, p + ".norm"), sd[p + ".reduction.weight"])


def patch_split(x, sd, p):
    """Linear(C->2C, no bias) -> PixelShuffle(2) (-> C/2 channels) -> LN(C/2)."""
    u = F.linear(x, sd[p + ".reduction.weight"]).permute(0, 3, 1, 2)
    return _ln(F.pixel_shuffle(u, 2).permute(0, 2, 3, 1), sd, p + ".norm")


def g_a(x, sd, embed_dim=48, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), ws=4):
    t = F.conv2d(x, sd["g_a.proj.weight"], sd["g_a.proj.bias"], stride=2).permute(0, 2, 3, 1)
    t = _ln(t, sd, "g_a.norm")
    for i, depth in enumerate(depths):
        dim = embed_dim * 2 ** i
        t = _stage(t, sd, f"g_a.layers.{i}", depth, dim // heads[i], ws)
        if i < len(depths) - 1:
            t = patch_merging(t, sd, f"g_a.layers.{i}.downsample")
    return t.permute(0, 3, 1, 2).contiguous()


def g_s(y_hat, sd, embed_dim=48, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), ws=4):
    n = len(depths)
    t = y_hat.permute(0, 2, 3, 1)
    for i in range(n):
        k = n - 1 - i
        dim = embed_dim * 2 ** k
        t = _stage(t, sd, f"g_s.layers.{i}", depths[k], dim // heads[k], ws)
        if i < n - 1:
            t = patch_split(t, sd, f"g_s.layers.{i}.upsample")
    t = t.permute(0, 3, 1, 2)
    t = F.pixel_shuffle(O.conv(t, sd, "g_s.end_conv.0.0"), 2)
    return O.conv(t, sd, "g_s.end_conv.1")


def _stack(x, sd, p):
    keys = sorted({int(k[len(p) + 1:].split(".")[0]) for k in sd if k.startswith(p + ".") and k.endswith(".weight")})
    t = x
    for j in keys[:-1]:
        t = O.gelu(O.conv(t, sd, f"{p}.{j}"))
    return O.conv(t, sd, f"{p}.{keys[-1]}")


@torch.no_grad()
def forward(sd, x, table=None, num_slices=12, max_support=6, collect=False):
    y = g_a(x, sd)
    t = y
    for j, s in ((0, 1), (2, 1), (4, 2), (6, 1)):
        t = O.gelu(O.conv(t, sd, f"h_a.{j}", s))
    z = O.conv(t, sd, "h_a.8", 2)
    z_hat, z_lik = O.entropy_bottleneck(z, sd)
    latent_scales = O.h_s(z_hat, sd, "h_scale_s")
    latent_means = O.h_s(z_hat, sd, "h_mean_s")
    hh, ww = y.shape[2:]
    y_hat_slices, liks, syms, idxs = [], [], [], []
    for i, y_i in enumerate(y.chunk(num_slices, 1)):
        support = y_hat_slices[:max_support]
        mean_support = torch.cat([latent_means] + support, 1)
        mu = _stack(mean_support, sd, f"cc_mean_transforms.{i}")[:, :, :hh, :ww]
        scale = _stack(torch.cat([latent_scales] + support, 1), sd, f"cc_scale_transforms.{i}")[:, :, :hh, :ww]
        y_hat_i = torch.round(y_i - mu) + mu
        liks.append(O.gaussian_likelihood(y_hat_i, scale, mu))
        if table is not None:
            syms.append(O.gc_symbols(y_i, mu))
            idxs.append(O.gc_indexes(scale, table))
        lrp = _stack(torch.cat([mean_support, y_hat_i], 1), sd, f"lrp_transforms.{i}")
        y_hat_slices.append(y_hat_i + 0.5 * torch.tanh(lrp))
    y_hat = torch.cat(y_hat_slices, 1)
    out = {"x_hat": g_s(y_hat, sd), "likelihoods": {"y": torch.cat(liks, 1), "z": z_lik}}
    if table is not None:
        out["symbols"], out["indexes"] = torch.cat(syms, 1), torch.cat(idxs, 1)
    if collect:
        out.update(y=y, z=z, y_hat=y_hat, latent_means=latent_means, latent_scales=latent_scales)
    return out
