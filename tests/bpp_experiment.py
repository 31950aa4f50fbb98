import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import resdsic_b200
from oracle import weights, wacnn_oracle as O
from tests.helpers import bpp_of
torch.set_num_threads(os.cpu_count())
sd = weights.make_state_dict(0, profile="lowrate")
x = weights.make_image(2, 512, 768, seed=3)
cache = "gpurun_out/ref_lowrate_kodak.npz"
ref = O.forward(sd, x, weights.scale_table())
m = resdsic_b200.WACNN().eval(); m.load_state_dict(sd, strict=True); m = m.to("cuda:0").set_precision("bf16")
r = m.symbols_and_indexes(x.to("cuda:0"))
xh = r["x_hat"].cpu(); ly = r["likelihoods"]["y"].cpu().numpy(); lz = r["likelihoods"]["z"].cpu().numpy()
print("lib", os.environ.get("RDSIC_LIB_PATH", "default"), "x_hat max", float((xh - ref["x_hat"]).abs().max()),
      "flips", float((r["y_symbols"].cpu() != ref["symbols"]).float().mean()))
for b in range(2):
    n = 512 * 768
    b0 = bpp_of(ref["likelihoods"]["y"][b].numpy(), ref["likelihoods"]["z"][b].numpy(), n); b1 = bpp_of(ly[b], lz[b], n)
    ps = lambda a: float(-10 * torch.log10(((a - x[b]) ** 2).mean()))
    print("  img", b, "bpp", b1, "ref", b0, "rel %.5f%%" % (100 * abs(b1 - b0) / b0), "dPSNR", abs(ps(xh[b]) - ps(ref["x_hat"][b])))
