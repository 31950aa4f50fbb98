"""TEST INFRASTRUCTURE -- a CPU interpreter of `rdsic_op` descriptors.

Executes a `resdsic_b200.program.Program` that was *built* over CPU tensors
(Ctx(build_only=True)) by following the semantics written in
include/resdsic_b200.h, using torch CPU ops.  It lets the `-m "not gpu"` suite
check the HOST-side graph builder (weight packing, deconv phase mapping,
pixel-shuffle addressing, the slice-loop slot plan, ld/coff arithmetic) against
the reference's golden vectors without a GPU.  It is never imported by the
product; the CUDA kernels are checked against the same goldens with -m gpu.
"""
import ctypes

import torch
import torch.nn.functional as F

from resdsic_b200 import _lib

_TORCH_DT = {_lib.F32: torch.float32, _lib.BF16: torch.bfloat16}


class Sim:
    def __init__(self, prog):
        self.prog = prog
        self.bases = []
        for t in prog.keep:
            if t is not None:
                self.bases.append((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), t))

    def _flat(self, ptr):
        """Flat typed tensor aliasing the storage starting at `ptr`."""
        for lo, hi, t in self.bases:
            if lo <= ptr < hi:
                flat = t.view(-1)
                off = (ptr - lo) // t.element_size()
                return flat[off:]
        raise KeyError(f"pointer {ptr:#x} not owned by the program")

    def _nhwc(self, v, B, H, W, Cn, Cfull=None):
        """Strided [B,H,W,Cn] alias of a view (writes go through)."""
        flat = self._flat(v.ptr)
        if v.nchw:
            Cf = Cn if Cfull is None else Cfull
            return flat[: B * Cf * H * W].view(B, Cf, H, W).permute(0, 2, 3, 1)[..., :Cn]
        return torch.as_strided(flat, (B, H, W, Cn), (H * W * v.ld, W * v.ld, v.ld, 1), v.coff)

    def _nhwc_copy(self, v, B, H, W, Cn):
        """Copy-op addressing: NCHW views honour ld (channels of the underlying tensor) and coff."""
        flat = self._flat(v.ptr)
        if v.nchw:
            Cf = v.ld if v.ld > 0 else Cn
            return flat[: B * Cf * H * W].view(B, Cf, H, W).permute(0, 2, 3, 1)[..., v.coff:v.coff + Cn]
        return torch.as_strided(flat, (B, H, W, Cn), (H * W * v.ld, W * v.ld, v.ld, 1), v.coff)

    # ------------------------------------------------------------------ ops
    def conv(self, d):
        G = d.groups if d.groups > 1 else 1
        # grouped form: group g reads Cin channels starting g * in_group_stride channels into the view
        x = torch.cat([self._nhwc(d.in_, d.B, d.H, d.W, d.in_group_stride * g + d.Cin)[..., d.in_group_stride * g:].float()
                       for g in range(G)], dim=-1)
        if d.a_square:
            x = x * x
        wflat = self._flat(d.weight)[: d.Cout * d.KH * d.KW * d.Cin].float()  # (rows may be padded past Cout)
        w = wflat.view(d.Cout, d.KH, d.KW, d.Cin).permute(0, 3, 1, 2).contiguous()
        bias = self._flat(d.bias)[: d.Cout] if d.bias else None
        need_h = (d.OH - 1) * d.stride + d.KH - d.pad_h
        need_w = (d.OW - 1) * d.stride + d.KW - d.pad_w
        # (plain NCHW-contiguous operands everywhere: the CPU convolution's summation order depends on the memory format, and
        # a grouped launch must reproduce the ungrouped launches of the same layers bit for bit, as the GPU kernel does)
        xp = F.pad(x.permute(0, 3, 1, 2), (d.pad_w, max(0, need_w - d.W), d.pad_h, max(0, need_h - d.H))).contiguous()
        if G == 1:
            v = F.conv2d(xp, w, bias, stride=d.stride)
        else:
            # a group is an n tile of the same GEMM: evaluate each as its own convolution, so that the values equal those of
            # the ungrouped launches of the same layers bit for bit (F.conv2d(groups=G) may sum in another order)
            Cg = d.Cout // G
            v = torch.cat([F.conv2d(xp[:, g * d.Cin:(g + 1) * d.Cin].contiguous(), w[g * Cg:(g + 1) * Cg].contiguous(),
                                    bias[g * Cg:(g + 1) * Cg] if bias is not None else None, stride=d.stride) for g in range(G)], 1)
        v = v[:, :, : d.OH, : d.OW]  # [B,Cout,OH,OW]
        assert v.shape[2:] == (d.OH, d.OW), (v.shape, d.OH, d.OW)
        if d.pixel_shuffle == 3:  # columns packed sub-position-major (s * C/4 + c): back to nn.PixelShuffle's 4 c + s
            Cq = d.Cout // 4
            v = v.reshape(v.shape[0], 4, Cq, *v.shape[2:]).transpose(1, 2).reshape(v.shape)
        if d.pixel_shuffle:
            v = F.pixel_shuffle(v, 2)
            Cv, sy, sx, oy, ox = d.Cout // 4, 1, 1, 0, 0
        else:
            Cv, sy, sx, oy, ox = d.Cout, d.osy, d.osx, d.ooy, d.oox
        v = v.permute(0, 2, 3, 1)  # NHWC on the (possibly shuffled) GEMM grid

        def sub(view):
            return self._nhwc(view, d.B, d.OHt, d.OWt, Cv)[:, oy::sy, ox::sx][:, : v.shape[1], : v.shape[2]]

        if d.tail_mode in (1, 2):  # fused GDN: x (bf16-rounded) * (r)sqrt(beta + gamma @ bf16(x^2))
            Cg = d.Cout
            gam = self._flat(d.tail_weight)[: Cg * Cg].float().view(Cg, Cg)
            beta = self._flat(d.tail_bias)[:Cg]
            x2 = (v * v).bfloat16().float()
            norm = x2 @ gam.t() + beta
            v = v.bfloat16().float() * (torch.sqrt(norm) if d.tail_mode == 2 else torch.rsqrt(norm))
        elif d.tail_mode == 3:  # fused ResidualUnit tail: gelu(W3 @ bf16(gelu(x)) + b3 + res)
            w3 = self._flat(d.tail_weight)[: d.tail_n * d.Cout].float().view(d.tail_n, d.Cout)
            b3 = self._flat(d.tail_bias)[: d.tail_n]
            v = F.gelu(v).bfloat16().float() @ w3.t() + b3
            Cv = d.tail_n
            v = F.gelu(v + sub(d.res).float())
        e = d.epilogue
        res = sub(d.res).float() if d.res.ptr else None
        aux = sub(d.aux).float() if d.aux.ptr else None
        if e == _lib.EPI_GELU:
            v = F.gelu(v)
        elif e == _lib.EPI_RES_GELU:
            v = F.gelu(v + res)
        elif e == _lib.EPI_ADD_RES:
            v = v + res
        elif e == _lib.EPI_GATE:
            v = aux * torch.sigmoid(v) + res
        elif e == _lib.EPI_GDN:
            v = res * torch.rsqrt(v)
        elif e == _lib.EPI_IGDN:
            v = res * torch.sqrt(v)
        elif e == _lib.EPI_LRP:
            v = res + 0.5 * torch.tanh(v)
        for view, sq in ((d.out, False), (d.out2, bool(d.out2_square)), (d.out3, False)):
            if view.ptr:
                dst = sub(view)
                dst.copy_((v * v if sq else v).to(dst.dtype))

    def attn(self, d):
        C, heads, ws, s = d.C, d.heads, d.ws, d.shift
        dh = C // heads
        qkv = self._nhwc(d.qkv, d.B, d.H, d.W, 3 * C).float()
        out = self._nhwc(d.out, d.B, d.H, d.W, C)
        table = self._flat(d.bias_table)[: (2 * ws - 1) ** 2 * heads].view(-1, heads)
        t = torch.arange(ws * ws)
        hi, wi = t // ws, t % ws
        bias = table[(hi[:, None] - hi[None] + ws - 1) * (2 * ws - 1) + (wi[:, None] - wi[None] + ws - 1)]  # [N,N,heads]
        for wh in range(d.H // ws):
            for ww in range(d.W // ws):
                hy, wx = wh * ws + hi, ww * ws + wi  # shifted-frame coords of the tokens
                oy, ox = (hy + s) % d.H, (wx + s) % d.W
                tok = qkv[:, oy, ox]  # [B,N,3C]
                q = tok[..., :C].reshape(d.B, -1, heads, dh).transpose(1, 2) * d.scale
                k = tok[..., C:2 * C].reshape(d.B, -1, heads, dh).transpose(1, 2)
                v = tok[..., 2 * C:].reshape(d.B, -1, heads, dh).transpose(1, 2)
                a = q @ k.transpose(-1, -2) + bias.permute(2, 0, 1)
                if s > 0:
                    rid = 3 * ((hy >= d.H - ws).long() + (hy >= d.H - s).long()) + (wx >= d.W - ws).long() + (wx >= d.W - s).long()
                    a = a + torch.where(rid[:, None] != rid[None, :], -100.0, 0.0)
                o = (torch.softmax(a, -1) @ v).transpose(1, 2).reshape(d.B, -1, C)
                out[:, oy, ox] = o.to(out.dtype)

    def eb(self, d):
        z = self._nhwc(d.z, d.B, d.h, d.w, d.C).float()
        P = self._flat(d.params)[: d.C * _lib.EB_STRIDE].view(d.C, _lib.EB_STRIDE)
        med = P[:, 58]
        r = torch.round(z - med)
        q = r + med

        def logits(v):  # v [...,C]
            a = P[:, 0:3] * v[..., None] + P[:, 33:36]
            a = a + P[:, 46:49] * torch.tanh(a)
            for k in range(3):
                M = P[:, 3 + 9 * k: 12 + 9 * k].view(d.C, 3, 3)
                a = torch.einsum("cji,...ci->...cj", M, a) + P[:, 36 + 3 * k: 39 + 3 * k]
                a = a + P[:, 49 + 3 * k: 52 + 3 * k] * torch.tanh(a)
            return (P[:, 30:33] * a).sum(-1) + P[:, 45]

        ql = q
        if d.noise.ptr:  # training mode: likelihood at z + noise
            ql = z + self._nhwc(d.noise, d.B, d.h, d.w, d.C).float()
            if d.noisy_out.ptr:
                self._nhwc(d.noisy_out, d.B, d.h, d.w, d.C).copy_(ql)
        lo, up = logits(ql - 0.5), logits(ql + 0.5)
        sg = -torch.sign(lo + up)
        lik = torch.abs(torch.sigmoid(sg * up) - torch.sigmoid(sg * lo)).clamp(min=d.lik_bound)
        self._nhwc(d.z_hat, d.B, d.h, d.w, d.C).copy_(q)
        n = d.B * d.C * d.h * d.w
        self._flat(d.lik)[:n].view(d.B, d.C, d.h, d.w).copy_(lik.permute(0, 3, 1, 2))
        if d.symbols:
            self._flat(d.symbols)[:n].view(d.B, d.C, d.h, d.w).copy_(r.permute(0, 3, 1, 2).int())

    def gc(self, d):
        y = self._nhwc(d.y, d.B, d.h, d.w, d.Cs).float()
        mu = self._nhwc(d.mu, d.B, d.h, d.w, d.Cs).float()
        sc = self._nhwc(d.scale, d.B, d.h, d.w, d.Cs).float()
        r = torch.round(y - mu)
        if d.sym_in:  # decoder side: symbols come from the entropy decoder
            n_ = d.B * d.Ctot * d.h * d.w
            r = self._flat(d.sym_in)[:n_].view(d.B, d.Ctot, d.h, d.w)[:, d.lik_coff:d.lik_coff + d.Cs].permute(0, 2, 3, 1).float()
        rm = r
        if d.mask.ptr:  # ResDSIC progressive stream
            m = self._nhwc(d.mask, d.B, d.h, d.w, d.Cs).float()
            sc = sc * m
            rm = r * m
        if d.scale_eps != 0.0:  # cimd: scale * mask + 1e-7
            sc = sc + torch.tensor(d.scale_eps, dtype=torch.float32)
        yh = rm + mu
        yl = r + mu
        if d.noise.ptr:  # training mode: likelihood at y + noise
            yl = y + self._nhwc(d.noise, d.B, d.h, d.w, d.Cs).float()
            if d.noisy_out.ptr:
                self._nhwc(d.noisy_out, d.B, d.h, d.w, d.Cs).copy_(yl)
        v = torch.abs(yl - mu)
        s = torch.clamp(sc, min=torch.tensor(d.scale_bound))
        c = float(-(2 ** -0.5))
        lik = (0.5 * torch.erfc(c * ((0.5 - v) / s)) - 0.5 * torch.erfc(c * ((-0.5 - v) / s))).clamp(min=d.lik_bound)
        table = self._flat(d.table)[: d.n_table]
        idx = torch.searchsorted(table[:-1].contiguous(), s.contiguous(), right=False).int()
        for k in range(3):
            if d.y_hat[k].ptr:
                dst = self._nhwc(d.y_hat[k], d.B, d.h, d.w, d.Cs)
                dst.copy_(yh.to(dst.dtype))
        n = d.B * d.Ctot * d.h * d.w
        sl = slice(d.lik_coff, d.lik_coff + d.Cs)
        self._flat(d.lik)[:n].view(d.B, d.Ctot, d.h, d.w)[:, sl].copy_(lik.permute(0, 3, 1, 2))
        if d.symbols:
            self._flat(d.symbols)[:n].view(d.B, d.Ctot, d.h, d.w)[:, sl].copy_(rm.permute(0, 3, 1, 2).int())
        if d.indexes:
            self._flat(d.indexes)[:n].view(d.B, d.Ctot, d.h, d.w)[:, sl].copy_(idx.permute(0, 3, 1, 2))

    def mask(self, d):
        ins = [self._nhwc(d.in_[i], d.B, d.H, d.W, d.C).float() for i in range(d.n_in)]
        if d.mode == 1:
            m = torch.pow(torch.sigmoid(ins[0]), self._flat(d.gamma)[: d.C])
        else:
            m = torch.sigmoid(sum(torch.sigmoid(t) for t in ins))
        self._nhwc(d.out, d.B, d.H, d.W, d.C).copy_(torch.round(m))

    def copy(self, d):
        src = self._nhwc_copy(d.src, d.B, d.H, d.W, d.C).float()
        if d.op == 4:
            src = src + self._nhwc_copy(d.src2, d.B, d.H, d.W, d.C).float()
        if d.op == 1:
            src = F.gelu(src)
        if d.op == 2:
            src = src * src
        if d.op == 3:
            src = src.clamp(0, 1)
        dst = self._nhwc_copy(d.dst, d.B, d.H, d.W, d.C)
        dst.copy_(src.to(dst.dtype))

    def patch(self, d):
        x = self._nhwc(d.src, d.B, d.H, d.W, d.C).float().permute(0, 3, 1, 2)
        cols = F.unfold(x, (d.KH, d.KW), padding=d.pad, stride=d.stride)  # [B, C*KH*KW, L], channel-major
        cols = cols.view(d.B, d.C, d.KH * d.KW, d.OH, d.OW).permute(0, 3, 4, 2, 1).reshape(d.B, d.OH, d.OW, -1)
        dst = self._nhwc(d.dst, d.B, d.OH, d.OW, d.Kp)
        dst.zero_()
        dst[..., : cols.shape[-1]] = cols.to(dst.dtype)

    def ln(self, d):
        x = torch.as_strided(self._flat(d.in_.ptr), (d.rows, d.C), (d.in_.ld, 1), d.in_.coff).float()
        g, b = self._flat(d.gamma)[: d.C], self._flat(d.beta)[: d.C]
        out = torch.as_strided(self._flat(d.out.ptr), (d.rows, d.C), (d.out.ld, 1), d.out.coff)
        out.copy_(F.layer_norm(x, (d.C,), g, b, d.eps).to(out.dtype))

    def run(self):
        disp = {_lib.OP_CONV: ("conv", self.conv), _lib.OP_ATTN: ("attn", self.attn), _lib.OP_EB: ("eb", self.eb),
                _lib.OP_GC: ("gc", self.gc), _lib.OP_COPY: ("copy", self.copy), _lib.OP_LN: ("ln", self.ln), _lib.OP_PATCH: ("patch", self.patch),
                _lib.OP_MASK: ("mask", self.mask)}
        for op in self.prog.ops:
            if op.kind in _lib.SYNC_OPS:
                continue  # lanes are a scheduling hint: program order is always a valid execution
            name, fn = disp[op.kind]
            fn(getattr(op.u, name))


def run_on_cpu(prog):
    Sim(prog).run()
