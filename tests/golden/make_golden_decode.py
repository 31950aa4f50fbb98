"""Golden fixture for the DECODER-side slice loop: the UNMODIFIED reference
`WACNN.decompress` (cnn.py:296-342) with its two entropy-coder calls replaced by replay stubs --
`entropy_bottleneck.decompress` returns the z_hat of the encoder pass and `RansDecoder.decode_stream`
returns, slice by slice, the symbols the reference's compress loop produced (rANS itself,
`compressai.ans`, is C++ that is absent from /root/reference and out of scope).  Everything
else -- hyper-synthesis, context transforms, index build, dequantize, LRP, g_s, clamp -- is the
reference's own code.  The stub also records the CDF indexes the decoder asks for.

    python tests/golden/make_golden_decode.py      (build container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim, wacnn_oracle, weights  # noqa: E402
from tests.golden.make_golden import run_reference_model  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"decode_c64x128": (2, 64, 128)}


def main():
    torch.set_num_threads(8)
    sd = weights.make_state_dict(seed=0)
    net = ref_shim.reference_wacnn().eval()
    net.load_state_dict(sd, strict=True)
    table = weights.scale_table()
    cnn_mod = sys.modules[type(net).__module__]
    for case, (B, H, W) in CASES.items():
        x = weights.make_image(B, H, W, seed=5)
        enc = run_reference_model(net, x, table)  # reference forward + compress-side loop (symbols, indexes, z_hat)
        z_hat = torch.from_numpy(enc["z_hat"])
        symbols = torch.from_numpy(enc["symbols"])
        asked = []

        class ReplayDecoder:
            def __init__(self):
                self.i = 0

            def set_stream(self, stream):
                pass

            def decode_stream(self, indexes, cdf, cdf_lengths, offsets):
                asked.append(np.asarray(indexes, dtype=np.int32))
                rv = symbols[:, 32 * self.i:32 * self.i + 32].reshape(-1).tolist()
                self.i += 1
                return rv

        cnn_mod.RansDecoder = ReplayDecoder
        net.entropy_bottleneck.decompress = lambda strings, size: z_hat
        with torch.no_grad():
            out = net.decompress([[b""], [b""]], z_hat.shape[-2:])
        h, w = H // 16, W // 16
        idx = np.concatenate([a.reshape(B, 32, h, w) for a in asked], 1)
        assert np.array_equal(idx, enc["indexes"]), "decoder asked for different CDF indexes than the encoder used"
        chk = wacnn_oracle.decode(sd, z_hat, symbols, table)
        assert torch.equal(chk["indexes"], torch.from_numpy(idx))
        assert (chk["x_hat"] - out["x_hat"]).abs().max().item() < 1e-5
        path = os.path.join(HERE, f"wacnn_{case}.npz")
        np.savez_compressed(path, z_hat=enc["z_hat"], symbols=enc["symbols"], indexes=idx, x_hat_dec=out["x_hat"].numpy(),
                            x_hat_enc=enc["x_hat"], y_hat=enc["y_hat"], image_seed=np.array(5))
        print(case, "max |x_hat_dec - clamp(x_hat_enc)|",
              float(np.abs(out["x_hat"].numpy() - np.clip(enc["x_hat"], 0, 1)).max()), os.path.getsize(path))


if __name__ == "__main__":
    main()
