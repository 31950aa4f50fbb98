import os, sys
sys.path.insert(0, "/root/repo")
import torch
from oracle import weights
from resdsic_b200.models import WACNN
sd = weights.make_state_dict(seed=0)
m = WACNN.from_state_dict(sd).to("cuda:0").eval()
m.set_precision("bf16")
x = weights.make_image(16, 256, 256, seed=2).to("cuda:0")
outs = []
for i in range(4):
    o = m(x)
    outs.append({k: v.clone() for k, v in [("x_hat", o["x_hat"]), ("ly", o["likelihoods"]["y"]), ("lz", o["likelihoods"]["z"])]})
for i in range(1, 4):
    print(i, {k: float((outs[i][k] - outs[0][k]).abs().max()) for k in outs[0]}, {k: int((outs[i][k] != outs[0][k]).sum()) for k in outs[0]})
