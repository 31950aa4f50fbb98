"""GPU probe (not a test): time the first layer's im2col (patchify) at the bench shape.

    RDSIC_PATCH_FIRST=0 python tests/gpu_patch_bench.py     (0 = the table-driven tiled kernel, 1 = the specialised one)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resdsic_b200.program import TV, Program  # noqa: E402

DEV = torch.device("cuda:0")


def run(B, H, W, iters=30):
    x = torch.rand(B, 3, H, W, device=DEV)
    OH, OW, Kp = H // 2, W // 2, 80
    out = TV(torch.empty(B * OH * OW * Kp, device=DEV, dtype=torch.bfloat16), B, OH, OW, Kp)
    prog = Program(DEV)
    prog.patchify(TV.nchw_of(x), out, 5, 5, 2, 2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        prog.run()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        prog.run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    byts = x.numel() * 4 + out.t.numel() * 2
    med = ts[len(ts) // 2]
    print(f"first={os.environ.get('RDSIC_PATCH_FIRST', '1')} B{B} {H}x{W}: median {med:7.1f} us  min {ts[0]:7.1f} us  {byts / med / 1e6:5.2f} TB/s  "
          f"checksum {out.t.float().sum().item():.6e}")


if __name__ == "__main__":
    run(24, 512, 768)
