"""GPU probe (not a test): time the window-attention core at the bench shape for one heads-per-CTA setting.

    RDSIC_ATTN_HPC=4 python tests/gpu_attn_bench.py        (8 = whole window per CTA, the round-1/2 form)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resdsic_b200.program import TV, Program  # noqa: E402

DEV = torch.device("cuda:0")


def run(B, H, W, C, ws, shift, heads=8, iters=30):
    qkv = (torch.rand(B * H * W * 3 * C, device=DEV) - 0.5).bfloat16()
    out = torch.zeros(B * H * W * C, dtype=torch.bfloat16, device=DEV)
    table = torch.rand((2 * ws - 1) ** 2, heads, device=DEV)
    prog = Program(DEV)
    prog.attn(TV(qkv, B, H, W, 3 * C), TV(out, B, H, W, C), table, heads, ws, shift, (C // heads) ** -0.5)
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
    byts = B * H * W * C * 2 * 4
    med = ts[len(ts) // 2]
    print(f"hpc={os.environ.get('RDSIC_ATTN_HPC', 'auto'):>4} B{B} {H}x{W} C{C} w{ws} shift{shift}: median {med:7.1f} us  min {ts[0]:7.1f} us  "
          f"{byts / med / 1e6:6.2f} TB/s  checksum {out.float().abs().sum().item():.6e}")


if __name__ == "__main__":
    run(24, 128, 192, 192, 8, 4)
    run(24, 128, 192, 192, 8, 0)
    run(24, 32, 48, 320, 4, 2)
