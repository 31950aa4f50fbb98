"""ctypes binding of libresdsic_b200.so (C ABI: include/resdsic_b200.h).

There is NO fallback: if the CUDA library is missing or its ABI does not match
this mirror, importing/using the package raises immediately.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# RDSIC_LIB_PATH: developer knob to A/B-test an alternative build of the same library (kernel tuning)
LIB_PATH = os.environ.get("RDSIC_LIB_PATH") or os.path.join(_PKG, "lib", "libresdsic_b200.so")

F32, BF16 = 0, 1
EPI_NONE, EPI_GELU, EPI_RES_GELU, EPI_ADD_RES, EPI_GATE, EPI_GDN, EPI_IGDN, EPI_LRP = range(8)
OP_CONV, OP_ATTN, OP_EB, OP_GC, OP_COPY, OP_LN, OP_PATCH, OP_FORK, OP_JOIN, OP_RECORD, OP_WAIT, OP_MASK = range(12)
SYNC_OPS = (OP_FORK, OP_JOIN, OP_RECORD, OP_WAIT)
EB_STRIDE = 60


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("ld", C.c_int32), ("coff", C.c_int32), ("nchw", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("in_", View),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("w_dtype", C.c_int32),
        ("Cout", C.c_int32), ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32),
        ("pad_h", C.c_int32), ("pad_w", C.c_int32),
        ("OH", C.c_int32), ("OW", C.c_int32),
        ("OHt", C.c_int32), ("OWt", C.c_int32), ("osy", C.c_int32), ("osx", C.c_int32),
        ("ooy", C.c_int32), ("oox", C.c_int32),
        ("pixel_shuffle", C.c_int32), ("epilogue", C.c_int32), ("a_square", C.c_int32),
        ("out", View), ("res", View), ("aux", View), ("out2", View), ("out3", View),
        ("out2_square", C.c_int32), ("tail_mode", C.c_int32), ("tail_weight", C.c_void_p), ("tail_bias", C.c_void_p),
        ("tail_n", C.c_int32), ("groups", C.c_int32), ("in_group_stride", C.c_int32), ("pad_", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("qkv", View), ("out", View), ("bias_table", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("heads", C.c_int32), ("ws", C.c_int32), ("shift", C.c_int32), ("scale", C.c_float),
    ]


class EBDesc(C.Structure):
    _fields_ = [
        ("z", View), ("z_hat", View), ("lik", C.c_void_p), ("symbols", C.c_void_p), ("params", C.c_void_p),
        ("B", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("C", C.c_int32), ("lik_bound", C.c_float),
        ("pad_", C.c_int32), ("noise", View), ("noisy_out", View),
    ]


class GCDesc(C.Structure):
    _fields_ = [
        ("y", View), ("mu", View), ("scale", View), ("y_hat", View * 3),
        ("lik", C.c_void_p), ("symbols", C.c_void_p), ("indexes", C.c_void_p), ("table", C.c_void_p),
        ("n_table", C.c_int32),
        ("B", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("Cs", C.c_int32), ("Ctot", C.c_int32),
        ("lik_coff", C.c_int32), ("scale_bound", C.c_float), ("lik_bound", C.c_float),
        ("scale_eps", C.c_float), ("noise", View), ("noisy_out", View), ("sym_in", C.c_void_p), ("mask", View),
    ]


class MaskDesc(C.Structure):
    _fields_ = [("in_", View * 8), ("out", View), ("gamma", C.c_void_p), ("n_in", C.c_int32), ("mode", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32)]


class CopyDesc(C.Structure):
    _fields_ = [("src", View), ("dst", View), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("C", C.c_int32), ("op", C.c_int32), ("pad_", C.c_int32), ("src2", View)]


class LNDesc(C.Structure):
    _fields_ = [("in_", View), ("out", View), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("rows", C.c_int32), ("C", C.c_int32), ("eps", C.c_float)]


class PatchDesc(C.Structure):
    _fields_ = [("src", View), ("dst", View), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("OH", C.c_int32),
                ("OW", C.c_int32), ("Kp", C.c_int32), ("pad_", C.c_int32)]


class SyncDesc(C.Structure):
    _fields_ = [("src", C.c_int32), ("event", C.c_int32)]


MAX_LANES = 12
MAX_EVENTS = 128


class _OpUnion(C.Union):
    _fields_ = [("conv", ConvDesc), ("attn", AttnDesc), ("eb", EBDesc), ("gc", GCDesc), ("copy", CopyDesc),
                ("ln", LNDesc), ("patch", PatchDesc), ("sync", SyncDesc), ("mask", MaskDesc)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lane", C.c_int32), ("u", _OpUnion)]


EXPORTS = (
    "rdsic_abi_version", "rdsic_error_string", "rdsic_sizeof",
    "rdsic_conv_forward", "rdsic_attn_forward", "rdsic_eb_forward", "rdsic_gc_forward", "rdsic_eb_aux_loss",
    "rdsic_gc_cdf_sizes", "rdsic_gc_pmf", "rdsic_eb_cdf_sizes", "rdsic_eb_pmf", "rdsic_pmf_to_quantized_cdf",
    "rdsic_copy_forward", "rdsic_ln_forward", "rdsic_patch_forward", "rdsic_mask_forward", "rdsic_run_program",
    "rdsic_graph_create", "rdsic_graph_launch", "rdsic_graph_num_kernels", "rdsic_graph_destroy",
    # training: backward twins (include/resdsic_b200.h, "Training")
    "rdsic_conv_dgrad_f32", "rdsic_conv_wgrad_f32", "rdsic_pointwise_f32", "rdsic_pixel_shuffle_f32",
    "rdsic_attn_backward_f32", "rdsic_gc_backward", "rdsic_eb_backward", "rdsic_eb_aux_backward", "rdsic_reduce_f32",
    # host-facing image I/O (csrc/image_io.cu)
    "rdsic_image_u8_to_f32", "rdsic_image_f32_to_u8", "rdsic_rate_per_image", "rdsic_rate_workspace_doubles",
)
(PW_ADD, PW_GELU_FWD, PW_GELU_BWD, PW_GATE_FWD, PW_GATE_BWD, PW_GDN_FWD, PW_GDN_BWD, PW_SQUARE_FWD, PW_SQUARE_BWD,
 PW_LRP_FWD, PW_LRP_BWD, PW_RECIP_SCALE, PW_DIFF_SCALE, PW_SCALE, PW_MUL) = range(15)
RED_SUM, RED_SUM_LOG, RED_SSE = range(3)

_lib = None


def lib():
    """Load (once) and return the CUDA library; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"resdsic_b200: CUDA library not built ({LIB_PATH}). Run `python -m resdsic_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RuntimeError(f"resdsic_b200: {LIB_PATH} does not export {name}")
    L.rdsic_error_string.restype = C.c_char_p
    L.rdsic_error_string.argtypes = [C.c_int]
    L.rdsic_sizeof.argtypes = [C.c_int]
    for fn, T in (("rdsic_conv_forward", ConvDesc), ("rdsic_attn_forward", AttnDesc), ("rdsic_eb_forward", EBDesc),
                  ("rdsic_gc_forward", GCDesc), ("rdsic_copy_forward", CopyDesc), ("rdsic_ln_forward", LNDesc),
                  ("rdsic_patch_forward", PatchDesc), ("rdsic_mask_forward", MaskDesc)):
        getattr(L, fn).argtypes = [C.POINTER(T), C.c_void_p]
        getattr(L, fn).restype = C.c_int
    L.rdsic_eb_aux_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rdsic_eb_aux_loss.restype = C.c_int
    vp, i32 = C.c_void_p, C.c_int32
    for fn, args in (("rdsic_gc_cdf_sizes", [vp, i32, C.c_float, vp, vp, vp]), ("rdsic_gc_pmf", [vp, vp, i32, vp, i32, vp]),
                     ("rdsic_eb_cdf_sizes", [vp, i32, vp, vp, vp]), ("rdsic_eb_pmf", [vp, vp, vp, vp, i32, i32, vp, i32, vp]),
                     ("rdsic_pmf_to_quantized_cdf", [vp, i32, vp, i32, i32, vp, i32, vp, vp])):
        getattr(L, fn).argtypes = args
        getattr(L, fn).restype = C.c_int
    L.rdsic_run_program.argtypes = [C.POINTER(Op), C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.rdsic_graph_create.argtypes = [C.POINTER(Op), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.rdsic_graph_launch.argtypes = [C.c_void_p, C.c_void_p]
    L.rdsic_graph_num_kernels.argtypes = [C.c_void_p]
    L.rdsic_graph_destroy.argtypes = [C.c_void_p]
    L.rdsic_graph_destroy.restype = None
    pv = C.POINTER(View)
    for fn, args in (("rdsic_conv_dgrad_f32", [C.POINTER(ConvDesc), vp, vp]),
                     ("rdsic_conv_wgrad_f32", [C.POINTER(ConvDesc), vp, vp, vp]),
                     ("rdsic_pointwise_f32", [C.c_int, C.c_size_t, vp, vp, vp, vp, vp, C.c_float, vp]),
                     ("rdsic_pixel_shuffle_f32", [C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
                     ("rdsic_attn_backward_f32", [C.POINTER(AttnDesc), pv, pv, vp, vp]),
                     ("rdsic_gc_backward", [C.POINTER(GCDesc), vp, pv, pv, pv, pv, vp]),
                     ("rdsic_eb_backward", [C.POINTER(EBDesc), vp, pv, pv, vp, vp]),
                     ("rdsic_eb_aux_backward", [vp, vp, vp, C.c_int, C.c_float, vp, vp]),
                     ("rdsic_reduce_f32", [C.c_int, C.c_size_t, vp, vp, vp, vp])):
        getattr(L, fn).argtypes = args
        getattr(L, fn).restype = C.c_int
    for fn, args in (("rdsic_image_u8_to_f32", [vp, vp, C.c_size_t, vp]), ("rdsic_image_f32_to_u8", [vp, vp, C.c_size_t, vp]),
                     ("rdsic_rate_per_image", [vp, C.c_size_t, vp, C.c_size_t, C.c_int, vp, vp, vp]),
                     ("rdsic_rate_workspace_doubles", [C.c_int])):
        getattr(L, fn).argtypes = args
        getattr(L, fn).restype = C.c_int
    if L.rdsic_abi_version() != 8:
        raise RuntimeError("resdsic_b200: ABI version mismatch between the python host and the CUDA library")
    for what, T in enumerate((Op, ConvDesc, AttnDesc, EBDesc, GCDesc, CopyDesc, View, LNDesc, PatchDesc, MaskDesc)):
        if L.rdsic_sizeof(what) != C.sizeof(T):
            raise RuntimeError(f"resdsic_b200: struct mirror mismatch for {T.__name__}: "
                               f"C={L.rdsic_sizeof(what)} python={C.sizeof(T)}")
    _lib = L
    return L


def check(rc, what="resdsic_b200 call"):
    if rc != 0:
        msg = lib().rdsic_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")
