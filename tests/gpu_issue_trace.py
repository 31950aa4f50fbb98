"""Profiling aid: prints the MMA issuer's clock64 stamps of one conv launch (RDSIC_TC_DBG_TS=1).
Columns per k-iteration: barrier wait, MMA issue, commit, and the gap to the next iteration."""
import ctypes
import json
import os
import sys

os.environ["RDSIC_TC_DBG_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from resdsic_b200 import _lib  # noqa: E402
from resdsic_b200.layers import Conv2d, Ctx  # noqa: E402
from resdsic_b200.program import TV  # noqa: E402


def main():
    cin, cout, k, stride, B, H, W = json.loads(os.environ.get("CASE", "[192,192,5,2,8,256,384]"))
    dev = "cuda:0"
    conv = Conv2d(cin, cout, k, stride).to(dev).set_precision("bf16")
    ctx = Ctx(dev, "bf16")
    x = TV(torch.randn(B * H * W * cin, device=dev).bfloat16(), B, H, W, cin)
    conv.emit(ctx, x)
    ctx.prog.run()
    ctx.prog.run()
    ctx.prog.run()  # (stamps are overwritten by every launch; the last one is read)
    buf = (ctypes.c_longlong * 4096)()
    lib = _lib.lib()
    rc = lib.rdsic_debug_read_ts(buf, 4096)
    assert rc == 0, rc
    t = list(buf)
    prev_end = None
    verbose = os.environ.get("VERBOSE")
    n = 0
    tot = dict(gap=0, wait=0, mma=0)
    slow = []
    for i in range(1024):
        a, b, c, d = t[4 * i:4 * i + 4]
        if d == 0:
            break
        gap = a - prev_end if prev_end is not None else 0
        if verbose:
            print(f"k{i:03d} gap {gap:5d} issue01 {b - a:5d} wait {c - b:5d} issue23 {d - c:5d} total {d - (prev_end or a):5d}")
        tot["gap"] += gap
        tot["wait"] += c - b
        tot["mma"] += (b - a) + (d - c)
        if gap > 400 or c - b > 300:
            slow.append((i, gap, c - b))
        prev_end = d
        n += 1
    span = t[4 * (n - 1) + 3] - t[0]
    print(f"{n} k-iterations in {span} cycles = {span / n:.0f} per iteration; mean gap {tot['gap'] / n:.0f} "
          f"wait {tot['wait'] / n:.0f} issue {tot['mma'] / n:.0f}")
    print("slow (index, gap, wait):", slow[:60])


if __name__ == "__main__":
    main()
