"""Probe (run on the GPU box): the CTA-pair ResidualUnit kernel inside the whole model, eager and graph mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import resdsic_b200
from resdsic_b200.utils import synthetic

B, H, W = int(os.environ.get("PB", 2)), int(os.environ.get("PH", 256)), int(os.environ.get("PW", 256))
m = resdsic_b200.WACNN().eval()
m.load_state_dict(synthetic.refinit_state_dict(0))
m = m.to("cuda:0").set_precision("bf16")
x = synthetic.rand_image(B, H, W, seed=1).to("cuda:0")
for graph in (False, True):
    m.use_cuda_graph = graph
    m._plans.clear() if hasattr(m, "_plans") else None
    for i in range(3):
        out = m(x)
        torch.cuda.synchronize()
        print("graph" if graph else "eager", i, float(out["x_hat"].float().mean()), flush=True)
