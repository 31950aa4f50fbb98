"""CPU-only: the C-ABI library loads and exports every symbol include/resdsic_b200.h declares, the ctypes
struct mirrors match, and the N>1 host logic (sharding, max-over-ranks, shard determinism) works over gloo."""
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    from resdsic_b200 import _lib
    header = open(os.path.join(ROOT, "include", "resdsic_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char\*)\s+(rdsic_\w+)\s*\(", header, flags=re.M))
    assert {"rdsic_conv_forward", "rdsic_attn_forward", "rdsic_eb_forward", "rdsic_gc_forward", "rdsic_run_program",
            "rdsic_graph_create", "rdsic_patch_forward"} <= declared
    L = _lib.lib()  # loads the .so, checks ABI version + every struct size (no CUDA call)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) == declared


def test_missing_library_fails_loudly(monkeypatch):
    from resdsic_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libresdsic_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_shard_range_partitions():
    from resdsic_b200.utils import shard_range
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                cover += list(range(lo, hi))
                assert 0 <= hi - lo <= -(-n // world)
            assert cover == list(range(n))
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import resdsic_b200
    from oracle import weights
    from resdsic_b200.utils import gather_to_rank0, max_over_ranks, shard_range
    from tests.program_sim import run_on_cpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    B = 3  # ragged on purpose: ranks get 2 and 1 images
    x = weights.make_image(B, 64, 64, seed=5)
    lo, hi = shard_range(B, rank, world)
    model = resdsic_b200.WACNN().eval()
    model.load_state_dict(weights.make_state_dict(0), strict=True)
    p = model._build(hi - lo, 64, 64, "cpu", True, build_only=True)
    p.x.copy_(x[lo:hi])
    run_on_cpu(p.prog)  # host-side graph builder under test; the CUDA kernels are covered by -m gpu
    x_hat = gather_to_rank0(p.x_hat)
    sym = gather_to_rank0(p.symbols)
    t_max = max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        q.put((x_hat.numpy(), sym.numpy(), t_max))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_rank_result():
    """world_size 2 over gloo: each rank runs its batch shard; the gathered result equals the one-rank
    result image for image (no cross-image reduction anywhere on the path), timings reduce with MAX."""
    import resdsic_b200
    from oracle import weights
    from tests.program_sim import run_on_cpu
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    x_hat, sym, t_max = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == [2.0, 10.0]
    model = resdsic_b200.WACNN().eval()
    model.load_state_dict(weights.make_state_dict(0), strict=True)
    full = model._build(3, 64, 64, "cpu", True, build_only=True)
    full.x.copy_(weights.make_image(3, 64, 64, seed=5))
    run_on_cpu(full.prog)
    np.testing.assert_array_equal(sym, full.symbols.numpy())
    np.testing.assert_allclose(x_hat, full.x_hat.numpy(), rtol=0, atol=1e-6)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys,
    on the arm's metric / unit, no GPU needed.  (Smallest run: one warm-up at 256 x 256, one timed 2-image step.)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "2"], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["metric"] == "WACNN (-m cnn) forward images/s at 768x512" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "batch-48" in line["config"]["workload"]
