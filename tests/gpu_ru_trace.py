"""Profiling aid: per-tile timeline of the fused ResidualUnit-tail kernel (conv_gdn_bf16.cu, RDSIC_TC_DBG_TS=1):
how long issuer 0 and the first epilogue warp of CTA 0 spend in each phase / wait of every tile."""
import ctypes
import os
import sys

os.environ["RDSIC_TC_DBG_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from resdsic_b200 import _lib  # noqa: E402
from resdsic_b200.layers import Ctx  # noqa: E402
from resdsic_b200.layers.layers import ResidualUnit  # noqa: E402
from resdsic_b200.program import TV  # noqa: E402


def main():
    B, H, W, N = 8, 128, 192, 192
    dev = "cuda:0"
    ru = ResidualUnit(N).to(dev).set_precision("bf16")
    ctx = Ctx(dev, "bf16")
    x = TV(torch.randn(B * H * W * N, device=dev).bfloat16(), B, H, W, N)
    ru.emit(ctx, x)
    for _ in range(3):
        ctx.prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ctx.prog.run()
    e1.record()
    torch.cuda.synchronize()
    print("ResidualUnit (2 launches) ms:", e0.elapsed_time(e1) / 5)
    buf = (ctypes.c_longlong * 4096)()
    assert _lib.lib().rdsic_debug_read_ts(buf, 4096) == 0
    t = list(buf)
    rows = []
    for i in range(250):
        s = t[16 * i:16 * i + 16]
        if s[5] == 0 or s[12] == 0 or (i and s[0] < t[16 * (i - 1)]):
            break
        rows.append(s)
    print(f"{len(rows)} tiles of CTA 0")
    hdr = "tile  period | issuer: wait_acc1_empty gemm1 wait_p wait_acc2_empty gemm2 | epilogue: wait_acc1_full phase1 wait_acc2_full phase2"
    print(hdr)
    for i, s in enumerate(rows):
        period = s[0] - rows[i - 1][0] if i else 0
        print(f"{i:4d} {period:7d} | {s[1] - s[0]:6d} {s[2] - s[1]:6d} {s[3] - s[2]:6d} {s[4] - s[3]:6d} {s[5] - s[4]:6d} | "
              f"{s[9] - s[8]:6d} {s[10] - s[9]:6d} {s[11] - s[10]:6d} {s[12] - s[11]:6d}")
    print("k-loop of tile 3 (issuer 0's own k-iterations): index, gap since previous, barrier wait, issue")
    prev = None
    for k in range(64):
        a, b, c, p = t[2048 + 4 * k:2048 + 4 * k + 4]
        if not c:
            continue
        patch = f" (A patch wait {p - a:6d})" if p else ""
        print(f"  k{k:02d} gap {a - prev if prev else 0:6d} wait {b - a:6d} issue {c - b:6d}{patch}")
        prev = c


if __name__ == "__main__":
    main()

