"""Host-side graph builder: a forward pass is a flat `Program` of kernel
descriptors (include/resdsic_b200.h `rdsic_op`) over caller-owned device
buffers.  Built once per input shape, executed natively (one C call per
forward, or a captured CUDA graph).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32, AttnDesc, ConvDesc, CopyDesc, EBDesc, GCDesc, LNDesc, MaskDesc, Op, PatchDesc, View

_DT = {torch.float32: F32, torch.bfloat16: BF16}


class TV:
    """A channels-last (or NCHW) view of a device tensor: element (b,y,x,c) at
    ((b*H+y)*W+x)*ld + coff + c."""

    __slots__ = ("t", "B", "H", "W", "C", "ld", "coff", "nchw")

    def __init__(self, t, B, H, W, C_, ld=None, coff=0, nchw=False):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C_
        self.ld = C_ if ld is None else ld
        self.coff, self.nchw = coff, nchw

    @staticmethod
    def empty(B, H, W, C_, dtype, device, ld=None):
        ld = C_ if ld is None else ld
        return TV(torch.empty(B * H * W * ld, dtype=dtype, device=device), B, H, W, C_, ld)

    @staticmethod
    def nchw_of(t):
        B, C_, H, W = t.shape
        assert t.is_contiguous()
        return TV(t, B, H, W, C_, nchw=True)

    @staticmethod
    def nchw_channels(t, off, n, H=None, W=None, total=None):
        """Channels [off, off+n) of a contiguous NCHW tensor as an NCHW view (copy ops only: `ld` = channel count
        of the underlying tensor, `coff` = channel offset).  `H, W, total` may reinterpret the per-image memory
        as [total, H, W] -- e.g. the reference's plain `reshape` of y_base (scalable/single_decoder.py:226-230)."""
        B, C_, Ht, Wt = t.shape
        assert t.is_contiguous()
        H, W, total = (Ht if H is None else H), (Wt if W is None else W), (C_ if total is None else total)
        assert total * H * W == C_ * Ht * Wt and off + n <= total
        return TV(t, B, H, W, n, ld=total, coff=off, nchw=True)

    def channels(self, off, n):
        """Channel sub-range [off, off+n) of the same pixels (torch chunk / cat as addressing)."""
        assert not self.nchw
        return TV(self.t, self.B, self.H, self.W, n, self.ld, self.coff + off)

    def view(self):
        v = View()
        if self.t is not None:
            v.ptr = self.t.data_ptr()
            v.dtype = _DT[self.t.dtype]
        v.ld, v.coff, v.nchw = self.ld, self.coff, int(self.nchw)
        return v

    def to_nchw(self):
        """Materialise as a torch NCHW tensor (host-side convenience for tests/outputs)."""
        if self.nchw:
            return self.t.view(self.B, self.C, self.H, self.W)
        full = self.t.view(self.B, self.H, self.W, self.ld)[..., self.coff:self.coff + self.C]
        return full.permute(0, 3, 1, 2).contiguous()


_NULL = View()


def _ptr(t):
    return None if t is None else t.data_ptr()


class Program:
    def __init__(self, device):
        self.device = device
        self.ops = []
        self.keep = []  # tensors referenced by raw pointer
        self._arr = None
        self._graph = None
        self._graph_stream = None
        self._lane = 0
        self._n_events = 0

    # ------------------------------------------------------------- lanes
    def _append(self, op):
        op.lane = self._lane
        self.ops.append(op)
        self._arr = None

    def sync(self, waiter, src, kind=None):
        """Lane `waiter` waits for everything issued so far on lane `src`."""
        assert 0 <= waiter < _lib.MAX_LANES and 0 <= src < _lib.MAX_LANES and waiter != src
        op = Op()
        op.kind = _lib.OP_FORK if kind is None else kind
        op.u.sync.src = src
        op.lane = waiter
        self.ops.append(op)
        self._arr = None

    def record(self):
        """Mark the current position of the current lane; returns an event id for `wait`."""
        assert self._n_events < _lib.MAX_EVENTS
        op = Op()
        op.kind = _lib.OP_RECORD
        op.u.sync.event = self._n_events
        self._n_events += 1
        self._append(op)
        return op.u.sync.event

    def wait(self, event):
        """The current lane waits for the position marked by `record()`."""
        op = Op()
        op.kind = _lib.OP_WAIT
        op.u.sync.event = event
        self._append(op)

    def fork(self, lane=1, src=None):
        """Start a concurrent branch on `lane` (it first waits for the current lane, or `src`)."""
        self.sync(lane, self._lane if src is None else src, _lib.OP_FORK)

    def join(self, lane=1, into=None):
        self.sync(self._lane if into is None else into, lane, _lib.OP_JOIN)

    def side(self, lane=1):
        """`with prog.side(k):` -- ops added inside run on lane k."""
        prog = self

        class _Side:
            def __enter__(self_):
                self_.prev = prog._lane
                prog._lane = lane

            def __exit__(self_, *a):
                prog._lane = self_.prev

        return _Side()

    @property
    def num_kernels(self):
        return sum(1 for op in self.ops if op.kind not in _lib.SYNC_OPS)

    # ------------------------------------------------------------- builders
    def conv(self, x: TV, weight, bias, Cout, KH, KW, stride, pad_h, pad_w, out: TV, epilogue=_lib.EPI_NONE,
             res: TV = None, aux: TV = None, out2: TV = None, out3: TV = None, a_square=False, pixel_shuffle=0,
             OH=None, OW=None, osy=1, osx=1, ooy=0, oox=0, Cin=None, out2_square=False, gdn=None, tail=None,
             groups=1, in_group_stride=0):
        """`groups` > 1: grouped form (include/resdsic_b200.h): `x` views the FIRST group's Cin channels, group g's
        start `g * in_group_stride` channels further into the same buffer (0: shared input); `weight` / `bias` / `res` /
        `out` carry the groups' Cout / groups channels back to back."""
        d = ConvDesc()
        d.in_ = x.view()
        d.B, d.H, d.W, d.Cin = x.B, x.H, x.W, (x.C if Cin is None else Cin)
        d.weight, d.bias = weight.data_ptr(), _ptr(bias)
        d.w_dtype = _DT[weight.dtype]
        d.Cout, d.KH, d.KW, d.stride, d.pad_h, d.pad_w = Cout, KH, KW, stride, pad_h, pad_w
        d.OH = (x.H + 2 * pad_h - KH) // stride + 1 if OH is None else OH
        d.OW = (x.W + 2 * pad_w - KW) // stride + 1 if OW is None else OW
        d.OHt, d.OWt = out.H, out.W
        d.osy, d.osx, d.ooy, d.oox = osy, osx, ooy, oox
        d.pixel_shuffle, d.epilogue, d.a_square = pixel_shuffle, epilogue, int(a_square)
        d.out = out.view()
        d.out2_square = int(out2_square)
        d.groups, d.in_group_stride = (groups, in_group_stride) if groups > 1 else (0, 0)
        if gdn is not None:  # (gamma' packed bf16, beta' fp32, inverse)
            d.tail_weight, d.tail_bias, d.tail_mode, d.tail_n = gdn[0].data_ptr(), gdn[1].data_ptr(), 2 if gdn[2] else 1, Cout
            self.keep += [gdn[0], gdn[1]]
        if tail is not None:  # ResidualUnit tail: (W3 packed bf16 [N2][Cout], b3 fp32, N2)
            d.tail_weight, d.tail_bias, d.tail_mode, d.tail_n = tail[0].data_ptr(), tail[1].data_ptr(), 3, tail[2]
            self.keep += [tail[0], tail[1]]
        for name, tv in (("res", res), ("aux", aux), ("out2", out2), ("out3", out3)):
            setattr(d, name, tv.view() if tv is not None else _NULL)
        assert weight.numel() >= Cout * KH * KW * d.Cin and weight.shape[-1] == KH * KW * d.Cin, \
            (weight.shape, Cout, KH, KW, d.Cin)
        op = Op()
        op.kind = _lib.OP_CONV
        op.u.conv = d
        self._append(op)
        self.keep += [x.t, weight, bias, out.t] + [tv.t for tv in (res, aux, out2, out3) if tv is not None]
        return out

    def attn(self, qkv: TV, out: TV, bias_table, heads, ws, shift, scale):
        d = AttnDesc()
        d.qkv, d.out = qkv.view(), out.view()
        d.bias_table = bias_table.data_ptr()
        d.B, d.H, d.W, d.C = out.B, out.H, out.W, out.C
        d.heads, d.ws, d.shift, d.scale = heads, ws, shift, scale
        op = Op()
        op.kind = _lib.OP_ATTN
        op.u.attn = d
        self._append(op)
        self.keep += [qkv.t, out.t, bias_table]
        return out

    def eb(self, z: TV, z_hat: TV, lik, params, symbols=None, lik_bound=1e-9, noise: TV = None, noisy_out: TV = None):
        d = EBDesc()
        d.z, d.z_hat = z.view(), z_hat.view()
        d.noise = noise.view() if noise is not None else _NULL
        d.noisy_out = noisy_out.view() if noisy_out is not None else _NULL
        self.keep += [tv.t for tv in (noise, noisy_out) if tv is not None]
        d.lik, d.symbols, d.params = lik.data_ptr(), _ptr(symbols), params.data_ptr()
        d.B, d.h, d.w, d.C = z.B, z.H, z.W, z.C
        d.lik_bound = lik_bound
        op = Op()
        op.kind = _lib.OP_EB
        op.u.eb = d
        self._append(op)
        self.keep += [z.t, z_hat.t, lik, params, symbols]

    def gc(self, y: TV, mu: TV, scale: TV, y_hat_dsts, lik, lik_coff, Ctot, table, symbols=None, indexes=None,
           scale_bound=0.11, lik_bound=1e-9, noise: TV = None, noisy_out: TV = None, sym_in=None, mask: TV = None, scale_eps=0.0):
        d = GCDesc()
        d.scale_eps = scale_eps
        d.y, d.mu, d.scale = y.view(), mu.view(), scale.view()
        d.mask = mask.view() if mask is not None else _NULL
        if mask is not None:
            self.keep.append(mask.t)
        d.sym_in = _ptr(sym_in)
        self.keep.append(sym_in)
        d.noise = noise.view() if noise is not None else _NULL
        d.noisy_out = noisy_out.view() if noisy_out is not None else _NULL
        self.keep += [tv.t for tv in (noise, noisy_out) if tv is not None]
        for i, tv in enumerate(y_hat_dsts):
            d.y_hat[i] = tv.view()
        d.lik, d.symbols, d.indexes = lik.data_ptr(), _ptr(symbols), _ptr(indexes)
        d.table, d.n_table = table.data_ptr(), table.numel()
        d.B, d.h, d.w, d.Cs, d.Ctot, d.lik_coff = y.B, y.H, y.W, y.C, Ctot, lik_coff
        d.scale_bound, d.lik_bound = scale_bound, lik_bound
        op = Op()
        op.kind = _lib.OP_GC
        op.u.gc = d
        self._append(op)
        self.keep += [y.t, mu.t, scale.t, lik, table, symbols, indexes] + [tv.t for tv in y_hat_dsts]

    def copy(self, src: TV, dst: TV, op_code=0, src2: TV = None):
        d = CopyDesc()
        d.src, d.dst = src.view(), dst.view()
        d.src2 = src2.view() if src2 is not None else _NULL
        d.B, d.H, d.W, d.C, d.op = src.B, src.H, src.W, src.C, op_code
        assert (op_code == 4) == (src2 is not None)
        op = Op()
        op.kind = _lib.OP_COPY
        op.u.copy = d
        self._append(op)
        self.keep += [src.t, dst.t] + ([src2.t] if src2 is not None else [])
        return dst

    def mask(self, ins, out: TV, mode, gamma=None):
        """ResDSIC importance mask (layers/mask_layer.py:64-107 + the eval-mode round of apply_noise)."""
        d = MaskDesc()
        for i, tv in enumerate(ins):
            d.in_[i] = tv.view()
        d.out, d.gamma, d.n_in, d.mode = out.view(), _ptr(gamma), len(ins), mode
        d.B, d.H, d.W, d.C = out.B, out.H, out.W, out.C
        op = Op()
        op.kind = _lib.OP_MASK
        op.u.mask = d
        self._append(op)
        self.keep += [tv.t for tv in ins] + [out.t, gamma]
        return out

    def patchify(self, x: TV, out: TV, KH, KW, stride, pad):
        d = PatchDesc()
        d.src, d.dst = x.view(), out.view()
        d.B, d.H, d.W, d.C = x.B, x.H, x.W, x.C
        d.KH, d.KW, d.stride, d.pad, d.OH, d.OW, d.Kp = KH, KW, stride, pad, out.H, out.W, out.C
        op = Op()
        op.kind = _lib.OP_PATCH
        op.u.patch = d
        self._append(op)
        self.keep += [x.t, out.t]
        return out

    def layernorm(self, x: TV, out: TV, gamma, beta, eps=1e-5):
        d = LNDesc()
        d.in_, d.out = x.view(), out.view()
        d.gamma, d.beta = gamma.data_ptr(), beta.data_ptr()
        d.rows, d.C, d.eps = x.B * x.H * x.W, x.C, eps
        op = Op()
        op.kind = _lib.OP_LN
        op.u.ln = d
        self._append(op)
        self.keep += [x.t, out.t, gamma, beta]
        return out

    # ------------------------------------------------------------- execution
    def _array(self):
        if self._arr is None or len(self._arr) != len(self.ops):
            self._arr = (Op * len(self.ops))(*self.ops)
        return self._arr

    def _require_cuda(self):
        if torch.device(self.device).type != "cuda":
            raise RuntimeError("resdsic_b200 programs execute on CUDA devices only (no CPU fallback)")

    @property
    def num_launches(self):
        return self.num_kernels

    def run(self, stream=None):
        """Launch every op in order on `stream` (default: torch's current stream)."""
        self._require_cuda()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        n, bad = C.c_int(0), C.c_int(-1)
        with torch.cuda.device(self.device):
            rc = _lib.lib().rdsic_run_program(self._array(), len(self.ops), stream, C.byref(n), C.byref(bad))
        _lib.check(rc, f"program op #{bad.value} (kind {self.ops[bad.value].kind if bad.value >= 0 else '?'})")
        return n.value

    def run_graph(self, stream=None):
        """Replay the program as one CUDA graph (captured on first use)."""
        self._require_cuda()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        L = _lib.lib()
        with torch.cuda.device(self.device):
            if self._graph is None:
                side = torch.cuda.Stream(self.device)  # capture must not use the legacy default stream
                side.wait_stream(torch.cuda.current_stream(self.device))
                h = C.c_void_p()
                _lib.check(L.rdsic_graph_create(self._array(), len(self.ops), side.cuda_stream, C.byref(h)),
                           "graph capture")
                self._graph, self._graph_stream = h, side
            _lib.check(L.rdsic_graph_launch(self._graph, stream), "graph launch")
        return self.num_kernels

    def __del__(self):
        try:
            if self._graph is not None:
                _lib.lib().rdsic_graph_destroy(self._graph)
        except Exception:
            pass
