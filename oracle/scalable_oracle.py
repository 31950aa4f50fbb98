"""TEST INFRASTRUCTURE -- CPU restatement (torch fp32, functional) of the ResDSIC scalable models' forward
(reference models/WACNN/scalable/single_decoder.py:343-504 `scalable_icd`, multiple_decoder.py:116-250
`scalable_imd`, conditional_single_decoder.py:109-271 `conditional_scalable_icd`, conditional_multiple_decoder.py:
104-268 `conditional_scalable_imd`, independent.py:289-461 `ResWACNNIndependentEntropy`) and of `Mask.forward` +
eval-mode `apply_noise` (layers/mask_layer.py:32-113).  Not product code.
Pinned against the unmodified reference by tests/golden/scalable_*.npz (tests/golden/make_golden_scalable.py).
"""
import torch
import torch.nn.functional as F

from . import wacnn_oracle as O


def g_a_split(x, sd, p="g_a"):
    """split_ga (:196-202): y_base = g_a[:6](x) -- up to and including the third conv, before its GDN."""
    t = O.gdn(O.conv(x, sd, f"{p}.0", 2), sd, f"{p}.1")
    t = O.gdn(O.conv(t, sd, f"{p}.2", 2), sd, f"{p}.3")
    t = O.attention_block(t, sd, f"{p}.4", 8, 4)
    y_base = O.conv(t, sd, f"{p}.5", 2)
    t = O.gdn(y_base, sd, f"{p}.6")
    t = O.conv(t, sd, f"{p}.7", 2)
    return y_base, t


def g_a_progressive(xp, sd, p="g_a_progressive"):
    """:87-96 (no trailing attention block)."""
    _, t = g_a_split(xp, sd, p)
    return t


def hyper(y, sd, suffix=""):
    """h_a -> EntropyBottleneck -> ste_round z_hat -> h_scale_s / h_mean_s (:360-381)."""
    t = y
    pa = "h_a" + suffix
    t = O.gelu(O.conv(t, sd, pa + ".0"))
    t = O.gelu(O.conv(t, sd, pa + ".2"))
    t = O.gelu(O.conv(t, sd, pa + ".4", 2))
    t = O.gelu(O.conv(t, sd, pa + ".6"))
    z = O.conv(t, sd, pa + ".8", 2)
    z_hat, z_lik = O.entropy_bottleneck(z, sd, "entropy_bottleneck" + suffix)
    return z, z_hat, z_lik, O.h_s(z_hat, sd, "h_scale_s" + suffix), O.h_s(z_hat, sd, "h_mean_s" + suffix)


def mask(policy, levels, scale, scale_prog, pr, sd, p="masking"):
    """Mask.forward (layers/mask_layer.py:41-113) followed by apply_noise(mask, False) = round for the learnable
    policies (single_decoder.py:399-401)."""
    if policy == "two-levels":
        return torch.zeros_like(scale) if pr == 0 else torch.ones_like(scale)
    inp = torch.cat([scale, scale_prog], 1) if scale_prog is not None else None  # (None: only the constant cases below are reachable)
    if policy == "learnable-mask-gamma":
        if pr == 0:
            return torch.zeros_like(scale)
        if pr == levels - 1:
            return torch.ones_like(scale)
        imp = torch.sigmoid(F.conv2d(inp, sd[f"{p}.mask_conv.0.weight"], sd[f"{p}.mask_conv.0.bias"]))
        index_pr = int(levels - 1 - pr)
        g = torch.sum(torch.stack([sd[f"{p}.gamma"][j] for j in range(index_pr)]), dim=0)[None, :, None, None]
        g = torch.relu(g) + 1e-7
        return torch.round(torch.pow(imp, g))
    if policy == "learnable-mask-nested":
        if pr == 0:
            return torch.zeros_like(scale)
        if pr == 1:
            return torch.ones_like(scale)
        s = sum(torch.sigmoid(F.conv2d(inp, sd[f"{p}.mask_conv.{i}.0.weight"], sd[f"{p}.mask_conv.{i}.0.bias"])) for i in range(pr))
        return torch.round(torch.sigmoid(s))
    raise NotImplementedError(policy)


@torch.no_grad()
def forward(sd, x, qualities, policy, levels, lrp_prog=True, independent_lrp=False, multiple_decoder=False, table=None,
            joiner_policy=None, variant=None):
    """scalable_icd.forward / scalable_imd.forward for a list of quality INDICES.  With `table` also returns the
    int32 symbols / indexes of both streams for the LAST quality (what compress hands the coder).
    `joiner_policy` (not None): conditional_scalable_icd.forward (conditional_single_decoder.py:112-271) -- the mask is
    computed without the progressive scales (:163), the progressive likelihood is taken at the UNMASKED scale (:221) and
    the two reconstructions of a slice are merged by `merge` (:103-113).
    `variant="cimd"`: conditional_scalable_imd.forward -- the mask IS computed from both scales
    (conditional_multiple_decoder.py:158), the progressive likelihood is taken at scale * mask + 1e-7 (:210; compress,
    inherited from `icd`, indexes scale * mask) and there are two decoders (:236).
    `variant="ind"`: ResWACNNIndependentEntropy.forward -- the mask of extract_mask is never applied
    (independent.py:318-392), i.e. an all-ones mask for every quality != 0."""
    B, _, H, W = x.shape
    y_base, y = g_a_split(x, sd)
    y = O.attention_block(y, sd, "g_a.8", 4, 2)
    N = y_base.shape[1]
    xp = torch.cat([y_base.reshape(B, N // 64, y_base.shape[2] * 8, y_base.shape[3] * 8), x], 1)  # concatenate, :226-230
    y_prog = g_a_progressive(xp, sd)
    z, z_hat, z_lik, lat_s, lat_m = hyper(y, sd)
    z_p, z_hat_p, z_lik_p, lat_sp, lat_mp = hyper(y_prog, sd, "_prog")
    base = O.slice_loop(y, lat_m, lat_s, sd, table)
    hh, ww = y.shape[2:]
    x_hats, y_hats, liks_p, masks = [], [], [], {}
    out = {}
    for q in qualities:
        y_hat_q = base["y_hat"]
        if q != 0:
            if variant == "ind":
                m = torch.ones_like(lat_s)
            else:
                m = mask(policy, levels, lat_s, lat_sp if (joiner_policy is None or variant == "cimd") else None, q, sd)
            masks[q] = m
            ys, ms = y_prog.chunk(10, 1), m.chunk(10, 1)
            hat, lk, syms, idxs = [], [], [], []
            for i in range(10):
                sup = hat[:5]
                mean_sup = torch.cat([lat_mp] + sup, 1)
                mu = O.cc_stack(mean_sup, sd, f"cc_mean_transforms_prog.{i}")[:, :, :hh, :ww]
                sc = O.cc_stack(torch.cat([lat_sp] + sup, 1), sd, f"cc_scale_transforms_prog.{i}")[:, :, :hh, :ww]
                if joiner_policy is None or variant == "cimd":
                    sc = sc * ms[i]
                r = torch.round(ys[i] - mu)
                sc_lik = sc + 1e-7 if variant == "cimd" else sc
                lk.append(O.gaussian_likelihood(r + mu, sc_lik, mu))  # gaussian_conditional_prog(y, scale*mask, mu), :447
                yh = r * ms[i] + mu  # :451
                if table is not None:
                    syms.append((r * ms[i]).to(torch.int32))
                    idxs.append(O.gc_indexes(sc, table))
                if lrp_prog:
                    fam = "lrp_transforms_prog" if independent_lrp else "lrp_transforms"
                    yh = yh + 0.5 * torch.tanh(O.cc_stack(torch.cat([mean_sup, yh], 1), sd, f"{fam}.{i}"))
                hat.append(yh)
            liks_p.append(torch.cat(lk, 1))
            if joiner_policy in (None, "residual"):
                y_hat_q = base["y_hat"] + torch.cat(hat, 1)
            elif joiner_policy in ("concatenation", "cac"):
                y_hat_q = base["y_hat"]
            else:  # "conditional": joiner[i](cat(y_hat_slice, y_hat_prog_slice))
                mains = base["y_hat"].chunk(10, 1)
                y_hat_q = torch.cat([joiner(torch.cat([mains[i], hat[i]], 1), sd, f"joiner.{i}") for i in range(10)], 1)
            if table is not None:
                out["prog_symbols"], out["prog_indexes"] = torch.cat(syms, 1), torch.cat(idxs, 1)
        y_hats.append(y_hat_q)
        gs = "g_s" if not multiple_decoder else ("g_s.0" if q == 0 else "g_s.1")
        g_in = y_hat_q
        if variant == "cimd" and joiner_policy == "concatenation" and q != 0:
            g_in = torch.cat([y_hat_q, torch.cat(hat, 1)], 1)  # conditional_multiple_decoder.py:230: a 2M-wide g_s[1]
        x_hats.append(g_s(g_in, sd, gs))
    lik_y = base["y_likelihoods"].unsqueeze(0)
    out.update(x_hat=torch.stack(x_hats), y=torch.stack(y_hats), z_hat=z_hat, z_hat_prog=z_hat_p, masks=masks,
               likelihoods={"y": lik_y, "z": z_lik, "z_prog": z_lik_p,
                            "y_prog": torch.stack(liks_p) if liks_p else torch.ones_like(lik_y)},
               y_base=y_base, y_lat=y, y_prog=y_prog)
    if table is not None:
        out["symbols"], out["indexes"] = base["symbols"], base["indexes"]
    return out


def joiner(t, sd, p):
    """conditional_single_decoder.py:39-48"""
    t = O.gelu(O.conv(t, sd, p + ".0"))
    t = O.gelu(O.conv(t, sd, p + ".2"))
    return O.conv(t, sd, p + ".4")


def g_s(y_hat, sd, p="g_s"):
    t = O.attention_block(y_hat, sd, f"{p}.0", 4, 2)
    t = O.gdn(O.deconv(t, sd, f"{p}.1"), sd, f"{p}.2", inverse=True)
    t = O.gdn(O.deconv(t, sd, f"{p}.3"), sd, f"{p}.4", inverse=True)
    t = O.attention_block(t, sd, f"{p}.5", 8, 4)
    t = O.gdn(O.deconv(t, sd, f"{p}.6"), sd, f"{p}.7", inverse=True)
    return O.deconv(t, sd, f"{p}.8")
