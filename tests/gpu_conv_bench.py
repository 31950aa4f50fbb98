"""Stand-alone timing of representative implicit-GEMM launches at Kodak batch-8 sizes (run on the GPU box;
also the target of the ncu captures under profiles/).  Not a pytest file."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from resdsic_b200.layers import Conv2d, Ctx
from resdsic_b200.program import TV

DEV = "cuda:0"
CASES = [  # name, Cin, Cout, k, stride, B, H, W (input)
    ("1x1_qkv_192_576", 192, 576, 1, 1, 8, 128, 192),
    ("1x1_ru_192_96", 192, 96, 1, 1, 8, 128, 192),
    ("3x3_ru_96_96", 96, 96, 3, 1, 8, 128, 192),
    ("5x5s2_192_192", 192, 192, 5, 2, 8, 256, 384),
    ("3x3_cc_352_224", 352, 224, 3, 1, 8, 32, 48),
    ("3x3_cc_64_32", 64, 32, 3, 1, 8, 32, 48),
    ("1x1_rutail_96_192_resgelu", 96, 192, 1, 1, 8, 128, 192),
]


def main():
    global CASES
    if os.environ.get("RDSIC_BENCH_CASES"):  # e.g. '[["k1_1024_192",1024,192,1,1,8,128,192]]'
        CASES = [tuple(c) for c in json.loads(os.environ["RDSIC_BENCH_CASES"])]
    only = sys.argv[1] if len(sys.argv) > 1 else None
    reps = int(os.environ.get("REPS", "5"))
    res = []
    for name, cin, cout, k, s, B, H, W in CASES:
        if only and only != name:
            continue
        conv = Conv2d(cin, cout, k, s).to(DEV).set_precision("bf16")
        ctx = Ctx(DEV, "bf16")
        x = TV(torch.randn(B * H * W * cin, device=DEV).bfloat16(), B, H, W, cin)
        kw = {}
        if "resgelu" in name:
            from resdsic_b200 import _lib
            kw = dict(epilogue=_lib.EPI_RES_GELU, res=TV(torch.randn(B * H * W * cout, device=DEV).bfloat16(), B, H, W, cout))
        out = conv.emit(ctx, x, **kw)
        for _ in range(2):
            ctx.prog.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.prog.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flop = 2.0 * B * out.H * out.W * cout * k * k * cin
        kiters = k * k * -(-cin // 64)
        tiles = -(-(B * out.H * out.W) // 128) * max(1, -(-cout // 256))
        cyc = ms * 1e-3 * 1.9e9 / (max(1.0, tiles / 148.0) * kiters)
        res.append(dict(name=name, ms=round(ms, 4), tflops=round(flop / ms / 1e9, 1), cyc_per_kiter=round(cyc)))
        print(res[-1], flush=True)
    return res


if __name__ == "__main__":
    main()
