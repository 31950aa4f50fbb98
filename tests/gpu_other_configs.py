"""Informational timings of the other BASELINE.json configurations (they are parity-test cases, not bench lines):
config 2 (batch 16 x 256 x 256, bf16) and config 5 (2048 x 1365 padded to 2048 x 1408, forward + symbol / index
build).  Device-resident, CUDA-event timed over graph replays.  Not a pytest file."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import resdsic_b200  # noqa: E402
from resdsic_b200.utils import synthetic  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    m = resdsic_b200.WACNN.from_state_dict(synthetic.make_state_dict(seed=0)).to("cuda:0").eval().set_precision("bf16")
    x = synthetic.make_image(16, 256, 256, seed=1).cuda()
    ms = timed(lambda: m(x), 20)
    print(f"config 2: batch 16 x 256 x 256 bf16 forward: {ms:.3f} ms/step = {16 / ms * 1e3:.0f} images/s = "
          f"{16 * 256 * 256 / ms / 1e3:.0f} MP/s")
    x = synthetic.make_image(1, 1408, 2048, seed=2).cuda()
    ms = timed(lambda: m.symbols_and_indexes(x), 10)
    print(f"config 5: 1 x 1408 x 2048 bf16 forward + symbols/indexes: {ms:.3f} ms = {1408 * 2048 / ms / 1e3:.0f} MP/s")
    x = synthetic.make_image(4, 1408, 2048, seed=2).cuda()
    ms = timed(lambda: m.symbols_and_indexes(x), 5)
    print(f"config 5 x4: 4 x 1408 x 2048: {ms:.3f} ms = {4 * 1408 * 2048 / ms / 1e3:.0f} MP/s")


if __name__ == "__main__":
    main()
