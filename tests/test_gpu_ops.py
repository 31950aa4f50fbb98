"""-m gpu: op-level parity of the CUDA kernels (through the C ABI) against the
oracle and the reference's golden vectors, on identical inputs."""
import os

import numpy as np
import pytest
import torch

import resdsic_b200
from oracle import wacnn_oracle as O
from oracle import weights
from resdsic_b200.layers import Ctx
from tests.conftest import GOLDEN
from tests.golden.make_golden import op_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def model(synthetic_sd):
    m = resdsic_b200.WACNN().eval()
    m.load_state_dict(synthetic_sd, strict=True)
    return m.to(DEV)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ops.npz"))


def test_library_loaded_is_in_tree():
    L = resdsic_b200._lib.lib()
    assert os.path.realpath(resdsic_b200._lib.LIB_PATH).startswith(os.path.realpath(os.path.dirname(resdsic_b200.__file__)))
    assert L.rdsic_abi_version() == 8


def test_gaussian_conditional_bit_exact_integers(model, gold, scale_table):
    """symbols / indexes / y_hat bit-exact given identical (y, mu, scale); likelihood to fp32 erfc ulps."""
    i = op_inputs()
    gc = model.gaussian_conditional
    y, mu, sc = (i[k].to(DEV) for k in ("gc_y", "gc_mu", "gc_scale"))
    out, lik = gc(y, sc, mu)
    np.testing.assert_array_equal(out.cpu().numpy(), gold["gc_yhat"])
    np.testing.assert_array_equal(gc.quantize(y, "symbols", mu).cpu().numpy(), gold["gc_sym"])
    np.testing.assert_array_equal(gc.quantize(y, "dequantize", mu).cpu().numpy(), gold["gc_yhat"])
    np.testing.assert_array_equal(gc.build_indexes(sc).cpu().numpy(), gold["gc_idx"])
    # CUDA erfcf vs the CPU libm erfc: a few ulps each, amplified by the upper-lower cancellation
    np.testing.assert_allclose(lik.cpu().numpy(), gold["gc_lik"], rtol=3e-4, atol=5e-7)
    # vs the oracle on a second, larger seeded input incl. every table threshold
    y2 = weights.hash_symmetric("gpu.gc.y", (3, 32, 24, 40), 20.0)
    mu2 = weights.hash_symmetric("gpu.gc.mu", (3, 32, 24, 40), 5.0)
    s2 = weights.hash_uniform("gpu.gc.s", (3, 32, 24, 40)) ** 5 * 400 - 0.1
    s2.view(-1)[:64] = scale_table
    s2.view(-1)[64:128] = torch.nextafter(scale_table, torch.tensor(0.0))
    s2.view(-1)[128:192] = torch.nextafter(scale_table, torch.tensor(1e9))
    assert torch.equal(gc.build_indexes(s2.to(DEV)).cpu(), O.gc_indexes(s2, scale_table))
    assert torch.equal(gc.quantize(y2.to(DEV), "symbols", mu2.to(DEV)).cpu(), O.gc_symbols(y2, mu2))
    yh = torch.round(y2 - mu2) + mu2
    _, lik2 = gc(y2.to(DEV), s2.to(DEV), mu2.to(DEV))
    np.testing.assert_allclose(lik2.cpu().numpy(), O.gaussian_likelihood(yh, s2, mu2).numpy(), rtol=3e-4, atol=5e-7)


def test_entropy_bottleneck(model, gold):
    z = op_inputs()["eb_z"]
    zh, lik = model.entropy_bottleneck(z.to(DEV))
    np.testing.assert_array_equal(zh.cpu().numpy(), gold["eb_zhat"])
    np.testing.assert_allclose(lik.cpu().numpy(), gold["eb_lik"], rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("kind,cin,B,hw", [("deconv", 192, 4, (64, 96)), ("deconv", 320, 3, (50, 70)), ("conv", 192, 3, (128, 192)),
                                           ("deconv", 192, 1, (12, 20))])
def test_fused_conv_gdn_vs_torch(kind, cin, B, hw):
    """conv / deconv + GDN in ONE kernel against torch on bf16-rounded operands, with x and x^2 rounded to bf16 where the
    kernels round them.  The large cases run gdn_pair_tc_kernel (CTA pairs; a ragged one with an odd tile count and
    Cin = 320), the small one conv_gdn_tc_kernel; stride-2 conv + GDN and the four deconv phases + inverse GDN."""
    from resdsic_b200.layers import GDN, Conv2d, ConvTranspose2d, Sequential
    import torch.nn.functional as F
    C = 192
    inverse = kind == "deconv"
    first = ConvTranspose2d(cin, C) if inverse else Conv2d(cin, C, 5, 2)
    gdn = GDN(C, inverse=inverse)
    with torch.no_grad():
        first.weight.copy_(weights.hash_symmetric(f"cg.w{kind}{cin}", first.weight.shape, (6.0 / (cin * 25)) ** 0.5))
        first.bias.copy_(weights.hash_symmetric(f"cg.b{kind}{cin}", first.bias.shape, 0.1))
        gdn.gamma.copy_(gdn.gamma_reparam.init(0.02 * torch.eye(C) + 0.002 * weights.hash_symmetric(f"cg.g{kind}", (C, C), 1.0).abs()))
        gdn.beta.copy_(gdn.beta_reparam.init(1.0 + 0.3 * weights.hash_symmetric(f"cg.be{kind}", (C,), 1.0).abs()))
    seq = Sequential(first, gdn)
    x = weights.hash_symmetric(f"cg.x{kind}{cin}{hw}", (B, cin, *hw), 1.0)
    if inverse:
        y = F.conv_transpose2d(_bf(x), _bf(first.weight.detach()), first.bias.detach(), stride=2, padding=2, output_padding=1)
    else:
        y = F.conv2d(_bf(x), _bf(first.weight.detach()), first.bias.detach(), stride=2, padding=2)
    from resdsic_b200 import packing
    gam, beta = packing.pack_gdn(gdn.beta, gdn.gamma, torch.bfloat16)
    norm = F.conv2d(_bf(y * y), gam.float()[:C, :C, None, None], beta.float())
    ref = _bf(y) * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
    out = seq.to(DEV).set_precision("bf16")(x.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("name,mod,key", [
    ("attn8", lambda m: m.g_a[4].conv_b[0], "attn8_x"),
    ("attn4", lambda m: m.g_a[8].conv_b[0], "attn4_x"),
    ("gdn", lambda m: m.g_a[1], "gdn_x"),
    ("igdn", lambda m: m.g_s[2], "gdn_x"),
    ("deconv", lambda m: m.g_s[3], "deconv_x"),
    ("block8", lambda m: m.g_a[4], "block8_x"),
    ("ru", lambda m: m.g_a[4].conv_a[0], "block8_x"),
])
def test_layers_fp32_vs_reference_golden(model, gold, name, mod, key):
    model.set_precision("fp32")
    x = op_inputs()[key].to(DEV)
    out = mod(model)(x).cpu().numpy()
    np.testing.assert_allclose(out, gold[name], rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("cin,cout,k,s,hw", [(3, 192, 5, 2, (34, 50)), (192, 192, 5, 2, (17, 24)), (96, 96, 3, 1, (9, 7)),
                                             (320, 288, 3, 2, (8, 12)), (352, 224, 3, 1, (5, 6)), (192, 96, 1, 1, (8, 8))])
def test_conv_shapes_vs_oracle(cin, cout, k, s, hw):
    """Ragged sizes (M not a tile multiple, odd H/W, stride 2) against torch CPU conv."""
    from resdsic_b200.layers import Conv2d
    conv = Conv2d(cin, cout, k, s)
    with torch.no_grad():
        conv.weight.copy_(weights.hash_symmetric(f"t.w{cin}{cout}{k}", conv.weight.shape, (3.0 / (cin * k * k)) ** 0.5))
        conv.bias.copy_(weights.hash_symmetric(f"t.b{cin}{cout}{k}", conv.bias.shape, 0.1))
    x = weights.hash_symmetric(f"t.x{cin}{hw}", (2, cin, *hw), 1.0)
    ref = torch.nn.functional.conv2d(x, conv.weight, conv.bias, stride=s, padding=k // 2)
    out = conv.to(DEV)(x.to(DEV)).cpu()
    np.testing.assert_allclose(out.detach().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-4)


def test_subpel_and_gelu_vs_oracle(model, synthetic_sd):
    x = weights.hash_symmetric("t.subpel", (2, 192, 3, 5), 1.0)
    ref = torch.nn.functional.pixel_shuffle(O.conv(x, synthetic_sd, "h_mean_s.2.0"), 2)
    out = model.h_mean_s[2](x.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)
    from resdsic_b200.layers import GELU
    np.testing.assert_allclose(GELU()(x.to(DEV)).cpu().numpy(), torch.nn.functional.gelu(x).numpy(), rtol=1e-5, atol=1e-6)
    out = model.h_a(weights.hash_symmetric("t.ha", (1, 320, 8, 12), 1.0).to(DEV)).cpu()
    ref = O.h_a(weights.hash_symmetric("t.ha", (1, 320, 8, 12), 1.0), synthetic_sd)
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=2e-4, atol=2e-4)


def test_error_paths(model):
    with pytest.raises(ValueError):
        model.g_a[4].conv_b[0](torch.zeros(1, 192, 12, 16, device=DEV))  # not a multiple of the window
    with pytest.raises(ValueError):
        model.gaussian_conditional.quantize(torch.zeros(1, 32, 4, 4, device=DEV), "bogus")
    with pytest.raises(ValueError, match="multiple of 64"):
        model(torch.zeros(1, 3, 100, 64, device=DEV))


# ----------------------------------------------------------------------------- bf16 tensor-core path
def _bf(x):
    return x.bfloat16().float()


@pytest.mark.parametrize("cin,cout,k,s,B,hw", [
    (64, 64, 1, 1, 1, (8, 16)), (192, 96, 1, 1, 2, (9, 7)), (96, 96, 3, 1, 2, (16, 24)), (352, 224, 3, 1, 1, (8, 12)),
    (288, 256, 3, 2, 1, (8, 12)), (192, 192, 5, 2, 1, (32, 48)), (192, 320, 5, 2, 2, (18, 26)), (320, 960, 1, 1, 1, (8, 12)),
    (64, 32, 3, 1, 3, (5, 3)), (512, 224, 3, 1, 1, (32, 48))])
def test_tcgen05_conv_vs_oracle(cin, cout, k, s, B, hw):
    """tcgen05/TMA implicit GEMM vs torch CPU conv on the same bf16-rounded operands (fp32 accumulate both
    sides): the only difference left is the bf16 rounding of the stored output (rel 2^-8)."""
    from resdsic_b200.layers import Conv2d
    conv = Conv2d(cin, cout, k, s)
    with torch.no_grad():
        conv.weight.copy_(weights.hash_symmetric(f"tc.w{cin}{cout}{k}", conv.weight.shape, (3.0 / (cin * k * k)) ** 0.5))
        conv.bias.copy_(weights.hash_symmetric(f"tc.b{cin}{cout}{k}", conv.bias.shape, 0.1))
    x = weights.hash_symmetric(f"tc.x{cin}{hw}", (B, cin, *hw), 1.0)
    ref = torch.nn.functional.conv2d(_bf(x), _bf(conv.weight.detach()), conv.bias.detach(), stride=s, padding=k // 2)
    out = conv.to(DEV).set_precision("bf16")(x.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("N,B,hw", [(192, 2, (16, 24)), (192, 3, (8, 16)), (192, 1, (19, 13)), (192, 2, (64, 96)), (192, 5, (128, 192)),
                                    (320, 1, (8, 12))])
def test_fused_residual_unit_vs_torch(N, B, hw):
    """Fused ResidualUnit (1x1 head launch + conv3x3 -> GELU -> conv1x1 -> +x -> GELU in ONE kernel: the CTA-pair kernel
    ru_pair_tc_kernel for N = 192, conv_gdn_tc_kernel<3> otherwise) against torch on the same bf16-rounded weights and
    input, with the intermediate activations rounded to bf16 where the kernels round them.  Shapes: an even and an odd
    number of 128-row tiles (the odd one out of the last CTA pair), ragged edges, one tile per CTA, several tiles per persistent CTA (accumulator double-buffering, barrier phases), N = 320."""
    from resdsic_b200.layers import ResidualUnit
    import torch.nn.functional as F
    ru = ResidualUnit(N)
    with torch.no_grad():
        for i, (cin, k) in zip((0, 2, 4), ((N, 1), (N // 2, 3), (N // 2, 1))):
            ru.conv[i].weight.copy_(weights.hash_symmetric(f"ru.w{N}{i}", ru.conv[i].weight.shape, (3.0 / (cin * k * k)) ** 0.5))
            ru.conv[i].bias.copy_(weights.hash_symmetric(f"ru.b{N}{i}", ru.conv[i].bias.shape, 0.2))
    x = weights.hash_symmetric(f"ru.x{N}{hw}", (B, N, *hw), 1.5)
    g = lambda v: F.gelu(v)
    c = [ru.conv[i] for i in (0, 2, 4)]
    xb = _bf(x)
    t = _bf(g(F.conv2d(xb, _bf(c[0].weight.detach()), c[0].bias.detach())))
    t = _bf(g(F.conv2d(t, _bf(c[1].weight.detach()), c[1].bias.detach(), padding=1)))
    ref = g(F.conv2d(t, _bf(c[2].weight.detach()), c[2].bias.detach()) + xb)
    out = ru.to(DEV).set_precision("bf16")(x.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("G,cin,cout,shared,B,hw", [(2, 224, 176, False, 2, (16, 24)), (10, 160, 224, True, 1, (16, 24)),
                                                    (10, 64, 32, False, 3, (32, 48)), (5, 32, 224, False, 1, (8, 12)),
                                                    (2, 176, 128, False, 24, (32, 48))])
def test_grouped_conv_vs_torch(G, cin, cout, shared, B, hw):
    """Grouped form of the implicit GEMM (rdsic_conv_desc.groups: one launch for G convolutions of identical shape --
    the slice loop's cc_mean / cc_scale pairs and its five independent tail slices) against torch's grouped conv on
    the same bf16-rounded operands; `shared`: every group reads the same input channels (in_group_stride = 0).
    Covers partial last K blocks (Cin = 176, 160, 224: the next group's channels sit in the unused half of the TMA box),
    one and many tiles per CTA, PAIR / M2 / single-issuer tile modes."""
    from resdsic_b200.layers import Conv2d, Ctx
    from resdsic_b200.layers.conv import emit_grouped
    import torch.nn.functional as F
    convs = [Conv2d(cin, cout, 3, 1) for _ in range(G)]
    with torch.no_grad():
        for g, c in enumerate(convs):
            c.weight.copy_(weights.hash_symmetric(f"grp.w{G}{cin}{cout}{g}", c.weight.shape, (3.0 / (cin * 9)) ** 0.5))
            c.bias.copy_(weights.hash_symmetric(f"grp.b{G}{cin}{cout}{g}", c.bias.shape, 0.1))
        convs = [c.to(DEV).set_precision("bf16") for c in convs]
    Ct = cin if shared else G * cin
    x = weights.hash_symmetric(f"grp.x{G}{cin}{hw}", (B, Ct, *hw), 1.0)
    xin = _bf(x) if not shared else _bf(x).repeat(1, G, 1, 1)
    wcat = torch.cat([_bf(c.weight.detach().cpu()) for c in convs], 0)
    bcat = torch.cat([c.bias.detach().cpu() for c in convs], 0)
    ref = F.gelu(F.conv2d(xin, wcat, bcat, padding=1, groups=G))
    ctx = Ctx(torch.device(DEV), "bf16")
    xt = ctx.from_nchw(x.to(DEV), torch.bfloat16)
    holder = convs[0]
    out = emit_grouped(ctx, holder, "t", convs, xt.channels(0, cin), 0 if shared else cin, gelu=True, out_dtype=torch.float32)
    res = ctx.to_nchw(out)
    ctx.prog.run()
    torch.cuda.synchronize()
    np.testing.assert_allclose(res.cpu().numpy(), ref.numpy(), rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("cin,cout,B,hw", [(192, 192, 2, (6, 10)), (320, 192, 2, (6, 10)), (192, 3, 2, (6, 10)),
                                           (192, 3, 3, (128, 96)), (192, 3, 2, (122, 99))])
def test_tcgen05_deconv_vs_oracle(cin, cout, B, hw):
    """(the two large 192 -> 3 cases run halo_pair_tc_kernel -- the image head on CTA pairs with halo patches --, the
    second one with ragged tile edges and an odd number of tiles)"""
    from resdsic_b200.layers import ConvTranspose2d
    d = ConvTranspose2d(cin, cout)
    with torch.no_grad():
        d.weight.copy_(weights.hash_symmetric(f"tc.dw{cin}{cout}", d.weight.shape, (12.0 / (cin * 25)) ** 0.5))
        d.bias.copy_(weights.hash_symmetric(f"tc.db{cin}{cout}", d.bias.shape, 0.1))
    x = weights.hash_symmetric(f"tc.dx{cin}{hw}", (B, cin, *hw), 1.0)
    ref = torch.nn.functional.conv_transpose2d(_bf(x), _bf(d.weight.detach()), d.bias.detach(), stride=2, padding=2,
                                               output_padding=1)
    out = d.to(DEV).set_precision("bf16")(x.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=8e-3, atol=1.5e-2)


@pytest.mark.parametrize("name,mod,key", [
    ("attn8", lambda m: m.g_a[4].conv_b[0], "attn8_x"), ("attn4", lambda m: m.g_a[8].conv_b[0], "attn4_x"),
    ("gdn", lambda m: m.g_a[1], "gdn_x"), ("igdn", lambda m: m.g_s[2], "gdn_x"), ("block8", lambda m: m.g_a[4], "block8_x"),
])
def test_layers_bf16_vs_reference_golden(model, gold, name, mod, key):
    """bf16 mode against the fp32 reference: error budget = bf16 operand/activation rounding."""
    model.set_precision("bf16")
    try:
        out = mod(model)(op_inputs()[key].to(DEV)).cpu().numpy()
    finally:
        model.set_precision("fp32")
    ref = gold[name]
    err = np.abs(out - ref)
    assert err.max() <= 0.05 * np.abs(ref).max() and err.mean() <= 0.01 * np.abs(ref).max(), (err.max(), err.mean())


@pytest.mark.parametrize("C,ws,shift,B,H,W", [(192, 8, 4, 2, 16, 24), (320, 4, 2, 3, 8, 12), (192, 8, 0, 1, 8, 8)])
def test_attention_core_bf16_vs_torch(C, ws, shift, B, H, W):
    """The tensor-core window-attention kernel alone (no qkv/proj GEMM) against a torch fp32 evaluation of
    the same bf16 qkv: roll / partition / bias / mask(-100) / softmax / PV / reverse, incl. shift = 0."""
    from resdsic_b200.program import TV, Program
    heads, dh = 8, C // 8
    qkv = weights.hash_symmetric(f"att.qkv{C}{ws}{shift}", (B, H, W, 3 * C), 2.0).bfloat16()
    table = weights.hash_symmetric(f"att.tab{ws}", ((2 * ws - 1) ** 2, heads), 1.0)
    scale = dh ** -0.5
    # torch reference on the shifted / partitioned tokens
    q = qkv.float()
    hs, wsf = (torch.arange(H) + shift) % H, (torch.arange(W) + shift) % W
    sh = q[:, hs][:, :, wsf]
    nWh, nWw, N = H // ws, W // ws, ws * ws
    win = sh.reshape(B, nWh, ws, nWw, ws, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(B * nWh * nWw, N, 3, heads, dh)
    qq, kk, vv = (win[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    att = (qq * scale) @ kk.transpose(-1, -2)
    idx = weights.relative_position_index(ws).reshape(-1)
    att = att + table[idx].reshape(N, N, heads).permute(2, 0, 1)[None]
    if shift:
        rid = O.shift_region_id(H, W, ws, shift).reshape(nWh, ws, nWw, ws).permute(0, 2, 1, 3).reshape(nWh * nWw, N)
        mask = torch.where(rid[:, None, :] != rid[:, :, None], -100.0, 0.0)
        att = (att.reshape(B, nWh * nWw, heads, N, N) + mask[None, :, None]).reshape(-1, heads, N, N)
    o = (torch.softmax(att, -1) @ vv).transpose(1, 2).reshape(B, nWh, nWw, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    ref = torch.empty(B, H, W, C)
    ref[:, hs[:, None], wsf[None, :]] = o.reshape(B, H, W, C)
    # kernel
    prog = Program(torch.device(DEV))
    tq = TV(qkv.to(DEV).reshape(-1), B, H, W, 3 * C)
    to = TV(torch.zeros(B * H * W * C, dtype=torch.bfloat16, device=DEV), B, H, W, C)
    prog.attn(tq, to, table.to(DEV), heads, ws, shift, scale)
    prog.run()
    got = to.t.float().cpu().reshape(B, H, W, C)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-4), ("bf16", 6e-2)])
def test_swin_block_vs_reference_tcm_golden(gold, precision, tol):
    """A14 (stf Swin block: LayerNorm kernel + GEMMs + fused window attention) vs the reference's tcm.Block."""
    from resdsic_b200.layers import SwinBlock
    from tests.golden.make_golden import SWIN_CASES, swin_state_dict
    for name, (dim, hd, ws, typ, B, H, W) in SWIN_CASES.items():
        blk = SwinBlock(dim, dim, hd, ws, 0.0, typ).eval()
        blk.load_state_dict(swin_state_dict(name, dim, hd, ws), strict=True)
        blk = blk.to(DEV).set_precision(precision)
        x = weights.hash_symmetric(f"{name}.x", (B, H, W, dim), 1.5)
        out = blk.forward_nhwc(x.to(DEV)).cpu().numpy()
        err = np.abs(out - gold[name])
        assert err.max() <= tol * max(1.0, np.abs(gold[name]).max()), (name, precision, err.max())


@pytest.mark.parametrize("C,k,s,p,B,H,W", [(3, 5, 2, 2, 2, 40, 72), (3, 5, 2, 2, 1, 64, 328), (3, 2, 2, 0, 2, 32, 48),
                                            (4, 3, 1, 1, 1, 9, 70), (16, 3, 1, 1, 1, 8, 8)])
def test_patchify_matches_unfold(C, k, s, p, B, H, W):
    """im2col of the first layer (tiled kernel for NCHW fp32 images with few channels, generic kernel
    otherwise): exactly torch's unfold, tap-major / channel-minor, zero padded to Kp, rounded to bf16."""
    from resdsic_b200.program import Program, TV
    x = weights.hash_symmetric(f"patch.{C}.{k}.{H}.{W}", (B, C, H, W), 2.0).to(DEV)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    Kp = -(-(k * k * C) // 16) * 16
    prog = Program(DEV)
    out = TV(torch.full((B * OH * OW * Kp,), 7.0, device=DEV, dtype=torch.bfloat16), B, OH, OW, Kp)
    prog.patchify(TV.nchw_of(x), out, k, k, s, p)
    prog.run()
    torch.cuda.synchronize()
    got = out.t.view(B, OH, OW, Kp).float()
    cols = torch.nn.functional.unfold(x, k, padding=p, stride=s)            # [B, C*k*k, OH*OW], channel-major
    want = cols.view(B, C, k * k, OH, OW).permute(0, 3, 4, 2, 1).reshape(B, OH, OW, k * k * C)
    assert torch.equal(got[..., :k * k * C], want.bfloat16().float())
    assert (got[..., k * k * C:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("env", [
    {"RDSIC_GDN_HALO": "1", "RDSIC_TC_HALO": "1"},  # experimental halo modes (off by default: measured slower)
    {"RDSIC_GDN_HALO": "2"},                         # halo with column-shifted, swizzle-atom-aligned patch copies
    {"RDSIC_TC_M2": "2"},                            # M2 (256-row tiles) forced on every eligible layer, however small
    {"RDSIC_TC_M2": "1"},                            # small grids with the K-split of two accumulators
    {"RDSIC_TC_M2": "0"},                            # M2 off
    {"RDSIC_TC_M2_MINK": "1", "RDSIC_TC_M2": "2"},   # M2 also for the 1-3 k-iteration pointwise GEMMs
    {"RDSIC_RU_DBL": "0", "RDSIC_RU_PAIR": "0"},     # single-buffered fused ResidualUnit kernel
    {"RDSIC_RU_PAIR": "0"},                          # 1-CTA double-buffered fused ResidualUnit kernel (default: CTA pairs)
    {"RDSIC_GDN_PAIR": "0", "RDSIC_HALO_PAIR": "0"},  # 1-CTA conv + GDN kernels, generic image head
    {"RDSIC_PDL": "1"},                              # programmatic dependent launch on every forward kernel (default off: measured slower)
    {"RDSIC_TC_PAIR": "3"},                          # cta_group::2 CTA pairs on EVERY layer with two M tiles (default: wide long-K layers)
    {"RDSIC_TC_PAIR": "0"},                          # no CTA pairs (M2 / single-issuer tiles everywhere)
    {"RDSIC_TC_PAIR": "0", "RDSIC_TC_MC": "1"},      # 1-CTA tiles with the B stage multicast across a CTA pair
    {"RDSIC_TC_PAIR": "3", "RDSIC_TC_P2": "2"},      # CTA pairs with 256-row A boxes (two accumulators per pair tile) wherever they fit
    {"RDSIC_ATTN_PIPE": "1"},                        # persistent double-buffered attention core (default off: measured slower)
    {"RDSIC_ATTN_HPC": "4"},                         # attention: a window's heads split over two CTAs
    {"RDSIC_PATCH_FIRST": "0"},                      # first layer's im2col through the table-driven tiled kernel
], ids=lambda e: ",".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_kernel_mode_switches_in_subprocess(env):
    """Every kernel mode behind a tuning switch stays correct: the conv / deconv / fused-layer parity tests pass with
    it selected.  The switches are read once per process, hence the subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_ops.py", "-m", "gpu", "-q", "-x", "-k",
                        "tcgen05 or layers_bf16 or subpel or fused_ or grouped or attention or patchify"], cwd=root, env=dict(os.environ, **env), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
