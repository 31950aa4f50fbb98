"""Diagnostic probe for the tcgen05 conv kernel (run on the GPU box): prints the error of a ladder of
cases from trivial (identity 1x1 GEMM) to the real layer shapes, so that a descriptor / swizzle / phase
bug can be localised from one run.  Not a pytest file."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import weights
from resdsic_b200.layers import Conv2d, ConvTranspose2d, Ctx, GDN
from resdsic_b200 import _lib

DEV = "cuda:0"


def bf(x):
    return x.bfloat16().float()


def run(mod, x, **kw):
    mod = mod.to(DEV).set_precision("bf16")
    ctx = Ctx(DEV, "bf16")
    out = ctx.to_nchw(mod.emit(ctx, ctx.from_nchw(x.to(DEV)), **kw))
    ctx.prog.run()
    torch.cuda.synchronize()
    return out.cpu()


def report(name, out, ref):
    err = (out - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    bad = (err > 2e-2 * scale).float().mean().item()
    print(f"{name:34s} max_err={err.max().item():.4e} ref_max={scale:.3e} bad_frac={bad:.4f} "
          f"{'OK' if bad == 0 else 'MISMATCH'}", flush=True)
    if bad > 0:
        idx = (err > 2e-2 * scale).nonzero()[:6]
        for i in idx:
            i = tuple(i.tolist())
            print("    at", i, "got", out[i].item(), "want", ref[i].item())
    return bad == 0


def conv_case(name, cin, cout, k, s, B, H, W, identity=False):
    c = Conv2d(cin, cout, k, s)
    with torch.no_grad():
        if identity:
            c.weight.zero_()
            for i in range(min(cin, cout)):
                c.weight[i, i, k // 2, k // 2] = 1.0
            c.bias.zero_()
        else:
            c.weight.copy_(weights.hash_symmetric(f"p.w{name}", c.weight.shape, (3.0 / (cin * k * k)) ** 0.5))
            c.bias.copy_(weights.hash_symmetric(f"p.b{name}", c.bias.shape, 0.1))
    x = weights.hash_symmetric(f"p.x{name}", (B, cin, H, W), 1.0)
    ref = F.conv2d(bf(x), bf(c.weight.detach()), c.bias.detach(), stride=s, padding=k // 2)
    return report(name, run(c, x), ref)


def main():
    ok = True
    ok &= conv_case("id_1x1_c64_m128", 64, 64, 1, 1, 1, 8, 16, identity=True)
    ok &= conv_case("id_1x1_c128_m256", 128, 128, 1, 1, 1, 16, 16, identity=True)
    ok &= conv_case("rand_1x1_c64_n64", 64, 64, 1, 1, 1, 8, 16)
    ok &= conv_case("rand_1x1_c192_n96_ragged", 192, 96, 1, 1, 2, 9, 7)
    ok &= conv_case("id_3x3_c64", 64, 64, 3, 1, 1, 8, 16, identity=True)
    ok &= conv_case("id_3x3_c64_halo16x16", 64, 64, 3, 1, 1, 16, 16, identity=True)
    ok &= conv_case("rand_3x3_c64_halo16x8", 64, 64, 3, 1, 1, 16, 8)
    ok &= conv_case("rand_3x3_c352_n224_halo", 352, 224, 3, 1, 2, 32, 48)
    ok &= conv_case("rand_3x3_c96_n96", 96, 96, 3, 1, 2, 16, 24)
    ok &= conv_case("rand_3x3_c352_n224", 352, 224, 3, 1, 1, 8, 12)
    ok &= conv_case("rand_3x3_s2_c288_n256", 288, 256, 3, 2, 1, 8, 12)
    ok &= conv_case("rand_5x5_s2_c192_n192", 192, 192, 5, 2, 1, 32, 48)
    ok &= conv_case("rand_5x5_s2_c192_n320_odd", 192, 320, 5, 2, 2, 18, 26)
    ok &= conv_case("rand_1x1_c320_n960", 320, 960, 1, 1, 1, 8, 12)
    # transposed conv (4 phases) incl. the 3-channel NCHW head
    for cin, cout, dhw in ((192, 192, (16, 8)),):
        d = ConvTranspose2d(cin, cout)
        with torch.no_grad():
            d.weight.copy_(weights.hash_symmetric(f"p.dwh{cin}{cout}", d.weight.shape, (12.0 / (cin * 25)) ** 0.5))
            d.bias.copy_(weights.hash_symmetric(f"p.dbh{cin}{cout}", d.bias.shape, 0.1))
        x = weights.hash_symmetric(f"p.dxh{cin}", (2, cin, *dhw), 1.0)
        ref = F.conv_transpose2d(bf(x), bf(d.weight.detach()), d.bias.detach(), stride=2, padding=2, output_padding=1)
        ok &= report(f"deconv_{cin}_{cout}_halo16x8", run(d, x), ref)
    for cin, cout in ((192, 192), (320, 192), (192, 3)):
        d = ConvTranspose2d(cin, cout)
        with torch.no_grad():
            d.weight.copy_(weights.hash_symmetric(f"p.dw{cin}{cout}", d.weight.shape, (12.0 / (cin * 25)) ** 0.5))
            d.bias.copy_(weights.hash_symmetric(f"p.db{cin}{cout}", d.bias.shape, 0.1))
        x = weights.hash_symmetric(f"p.dx{cin}", (2, cin, 6, 10), 1.0)
        ref = F.conv_transpose2d(bf(x), bf(d.weight.detach()), d.bias.detach(), stride=2, padding=2, output_padding=1)
        ok &= report(f"deconv_{cin}_{cout}", run(d, x), ref)
    # GDN (x^2 side buffer path)
    g = GDN(192)
    x = weights.hash_symmetric("p.gdn", (1, 192, 8, 16), 2.0)
    from resdsic_b200 import packing
    gam, beta = packing.pack_gdn(g.beta, g.gamma)
    xb = bf(x)
    norm = F.conv2d(bf(xb * xb), bf(gam).reshape(192, 192, 1, 1), beta)
    ok &= report("gdn_192", run(g, x), xb * torch.rsqrt(norm))
    # fused conv -> GDN and deconv -> IGDN (one kernel each; x^2 goes through TMEM)
    from resdsic_b200.layers import Sequential
    for name, first, inverse, xin in (
            ("fused_conv5x5s2_3_192_gdn", Conv2d(3, 192, 5, 2), False, weights.hash_uniform("p.fx1", (2, 3, 40, 56))),
            ("fused_deconv_192_192_igdn", ConvTranspose2d(192, 192), True, weights.hash_symmetric("p.fx2", (2, 192, 9, 13), 1.0)),
            ("fused_conv1x1_64_128_gdn", Conv2d(64, 128, 1, 1), False, weights.hash_symmetric("p.fx3", (1, 64, 8, 16), 1.0))):
        gd = GDN(first.out_channels, inverse=inverse)
        with torch.no_grad():
            first.weight.copy_(weights.hash_symmetric(f"p.fw{name}", first.weight.shape, 0.08))
            first.bias.copy_(weights.hash_symmetric(f"p.fb{name}", first.bias.shape, 0.1))
            gd.gamma.add_(weights.hash_uniform(f"p.fg{name}", gd.gamma.shape) * 0.02)
        seq = Sequential(first, gd)
        ctx_probe = Ctx(DEV, "bf16")
        n_before = 0
        out = run(seq, xin)
        xb = bf(xin)
        if isinstance(first, ConvTranspose2d):
            t = F.conv_transpose2d(xb, bf(first.weight.detach().cpu()), first.bias.detach().cpu(), stride=2, padding=2, output_padding=1)
        else:
            t = F.conv2d(xb, bf(first.weight.detach().cpu()), first.bias.detach().cpu(), stride=first.stride, padding=first.padding)
        gam, beta = packing.pack_gdn(gd.beta.detach().cpu(), gd.gamma.detach().cpu())
        C = first.out_channels
        norm = F.conv2d(bf(t * t), bf(gam).reshape(C, C, 1, 1), beta)
        ref = bf(t) * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
        ok &= report(name, out, ref)
    # fused ResidualUnit tail: conv3x3 -> GELU -> conv1x1 -> +x -> GELU in one kernel
    from resdsic_b200.layers import ResidualUnit
    for N_, hw in ((192, (16, 24)), (320, (8, 12))):
        ru = ResidualUnit(N_)
        with torch.no_grad():
            for j, c in ((0, ru.conv[0]), (2, ru.conv[2]), (4, ru.conv[4])):
                c.weight.copy_(weights.hash_symmetric(f"p.ru{N_}.{j}", c.weight.shape, (3.0 / (c.in_channels * c.kernel_size ** 2)) ** 0.5))
                c.bias.copy_(weights.hash_symmetric(f"p.rub{N_}.{j}", c.bias.shape, 0.1))
        x = weights.hash_symmetric(f"p.rux{N_}", (2, N_, *hw), 1.0)
        out = run(ru, x)
        xb = bf(x)
        w = lambda c: bf(c.weight.detach().cpu())
        t = bf(F.gelu(F.conv2d(xb, w(ru.conv[0]), ru.conv[0].bias.detach().cpu())))
        t = bf(F.gelu(F.conv2d(t, w(ru.conv[2]), ru.conv[2].bias.detach().cpu(), padding=1)))
        ref = F.gelu(F.conv2d(t, w(ru.conv[4]), ru.conv[4].bias.detach().cpu()) + xb)
        ok &= report(f"fused_ru_{N_}", out, ref)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
