"""Probe: time of the fused ResidualUnit tail kernel alone (B = 24, 128 x 192 x 192) under the current env knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
TS = torch.zeros(60 * 16, dtype=torch.int64, device="cuda:0")
if os.environ.get("TRACE"):
    os.environ["RDSIC_RP_TS"] = str(TS.data_ptr())
from resdsic_b200.layers import Conv2d, Ctx
from resdsic_b200.layers.conv import conv1x1, conv3x3
from resdsic_b200.program import TV
DEV = "cuda:0"
B, H, W, N = 24, 128, 192, 192
c3, c1 = conv3x3(N // 2, N // 2).to(DEV).set_precision("bf16"), conv1x1(N // 2, N).to(DEV).set_precision("bf16")
ctx = Ctx(DEV, "bf16")
t = TV(torch.randn(B * H * W * N // 2, device=DEV).bfloat16(), B, H, W, N // 2)
x = TV(torch.randn(B * H * W * N, device=DEV).bfloat16(), B, H, W, N)
w3, b3 = c1.packed(torch.bfloat16)
out = c3.emit(ctx, t, tail=(w3, b3, N), res=x)
for _ in range(3):
    ctx.prog.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    ctx.prog.run()
e1.record()
torch.cuda.synchronize()
print({k: v for k, v in os.environ.items() if k.startswith("RDSIC_")}, "us per launch:", round(1000 * e0.elapsed_time(e1) / reps, 1), flush=True)

if os.environ.get("TRACE"):
    t = TS.cpu().numpy().reshape(60, 16)
    t0 = t[0][0]
    names = ["i0.top", "i0.go", "i0.done", "i1.top", "i1.go", "i1.done", "tail.go", "p1.top", "p1.go", "p1.end", "p2.top", "p2.go", "p2.end"]
    print("tile " + " ".join(f"{n:>8s}" for n in names))
    for lt in range(2, 14):
        print(f"{lt:4d} " + " ".join(f"{int(v - t0):8d}" for v in t[lt][:13]))
