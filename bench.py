#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): WACNN (`-m cnn`)
forward at Kodak size 768x512, images/s (and MP/s), weak-scaled over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward pass over one batch of synthetic images per GPU.
Prints ONE JSON line on rank 0 (see DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W = 512, 768                 # Kodak size (BASELINE.json configs[2])
FLOP_PER_IMAGE = 413.22e9       # 1.0509 MFLOP per input pixel (conv+linear+bmm, 2*MAC)
METRIC = "WACNN (-m cnn) forward images/s at 768x512"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def _run_nvml(self):
        """Fast path: NVML through nvidia_ml_py (a query takes microseconds, so even a 100 ms timed region gets
        several samples); same fields as the nvidia-smi query below."""
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self._stop_evt.is_set():
            r = get_reasons(h)
            self.samples.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx),
                                 str(N.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                                ["Active" if r & b else "Not Active" for _, b in bits])
            self._stop_evt.wait(0.01)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def build_inputs(batch, seed):
    from resdsic_b200.utils import synthetic  # hash-seeded synthetic images (data generator, no compute)
    return synthetic.make_image(batch, H, W, seed=seed)


def make_weights():
    from resdsic_b200.utils import synthetic
    return synthetic.make_state_dict(seed=0)


# ----------------------------------------------------------------------------- reference / CPU arm
def time_cpu_oracle(sd, steps, warmup, batch):
    """The reference's CPU implementation of the path: the oracle port (torch fp32 ATen ops, the same
    operator calls the reference module makes) on all host threads."""
    from oracle import wacnn_oracle as O
    x = build_inputs(batch, seed=11)
    small = build_inputs(1, seed=12)[:, :, :256, :256].contiguous()
    for _ in range(max(1, warmup)):
        O.forward(sd, small if warmup > 1 else x)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward(sd, x)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_weights()
    ips, per_step = time_cpu_oracle(sd, args.steps, args.warmup, batch=1)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cnn forward 512x768, 1 image per step on host CPU", "precision": "fp32"},
        "megapixels_per_s": ips * H * W / 1e6,
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x 1 image 512x768, oracle (torch CPU fp32 restatement of "
                                   "the reference forward; the Python reference itself cannot travel to the GPU box)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist

    import resdsic_b200
    from resdsic_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (see module docstring)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    _lib.lib()  # fail loudly if the CUDA library is missing
    sd = make_weights()
    if args.model == "stf":  # builder-defined model (no reference implementation): informational runs only
        from resdsic_b200.utils import synthetic as _w
        model = resdsic_b200.models["stf"]().eval()
        sd = _w.synth_state_dict(model.state_dict())
    else:
        model = resdsic_b200.WACNN().eval()
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).set_precision(args.precision)
    model.micro_batches = args.micro_batches if args.micro_batches == "auto" else int(args.micro_batches)
    B = args.batch
    x_host = build_inputs(B, seed=100 + rank).pin_memory()
    x_dev = x_host.to(dev)

    # ---- device-resident timing (value): inputs already in HBM
    for _ in range(max(3, args.warmup)):
        out = model(x_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = model(x_dev)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches_per_step = model.last_num_launches
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end timing (e2e): through the public host-facing API (resdsic_b200.utils.ForwardPipeline):
    #      every step copies its batch from pinned host memory to the device and its results (x_hat + both
    #      likelihood tensors) back to pinned host memory; copies of neighbouring steps overlap the compute.
    from resdsic_b200.utils import ForwardPipeline
    pipe = ForwardPipeline(model, x_host, depth=2)
    host_batches = [x_host, build_inputs(B, seed=200 + rank).pin_memory()]
    pipe.run([host_batches[i % 2] for i in range(3)])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run([host_batches[i % 2] for i in range(args.steps)])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d_bytes, d2h_bytes = pipe.h2d_bytes, pipe.d2h_bytes

    from resdsic_b200.utils import max_over_ranks
    ms, ms_e2e = max_over_ranks([ms, ms_e2e], device=dev)

    # ---- per-kernel-family device time of one step (eager, CUDA events around every launch)
    fam = None
    if rank == 0:
        fam = profile_families(model, x_dev, args.dump_ops)

    if rank == 0:
        peaks = load_peaks()
        n_img = B * world * args.steps
        ips = n_img / (ms / 1e3)
        ips_e2e = n_img / (ms_e2e / 1e3)
        conv_ms, conv_n = fam["conv"]
        total_ms = sum(v[0] for v in fam.values())
        flop_per_image = FLOP_PER_IMAGE
        if args.model != "cnn":  # no survey figure: sum 2*M*N*K over the program's GEMM descriptors
            plan = next(iter(model._plans.values()))
            flop_per_image = sum(2.0 * o.u.conv.B * o.u.conv.OH * o.u.conv.OW * o.u.conv.Cout * o.u.conv.KH * o.u.conv.KW *
                                 o.u.conv.Cin for o in plan.prog.ops if o.kind == _lib.OP_CONV) / plan.sub_batch
        # the profile covers ONE sub-batch program (the model runs `micro_batches` of them concurrently)
        b_prog = next(iter(model._plans.values())).sub_batch
        tf = b_prog * flop_per_image / conv_n / (conv_ms / conv_n / 1e3) / 1e12  # algorithmic FLOP per launch / avg launch time
        traffic_bytes, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")
        if os.path.exists(tpath):  # DRAM bytes of the heaviest launch, from the committed ncu --set full capture
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic_bytes = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_note = f"{tj['kernel']}: algorithmic {tj['algorithmic_bytes']} B; {tj['source']}"
        line = {
            "metric": METRIC if args.model == "cnn" else METRIC.replace("WACNN (-m cnn)", "STF (-m stf, builder-defined)"), "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": (f"cnn (WACNN N=192 M=320) forward 512x768, batch {B} per GPU, eval mode" if args.model == "cnn"
                                    else f"stf (builder-defined, N=192 M=384) forward 512x768, batch {B} per GPU, eval mode"),
                       "precision": args.precision, "batch_per_gpu": B, "image": [H, W], "parallelism": f"dp{world}",
                       "l2": f"per-step activation working set (~{0.19 * B:.1f} GB at batch {B}) exceeds the 126 MB L2; no explicit flush",
                       "cuda_graph": bool(model.use_cuda_graph),
                       "micro_batches": len(next(iter(model._plans.values())).subs)},
            "megapixels_per_s": ips * H * W / 1e6,
            "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "api": "resdsic_b200.utils.ForwardPipeline.run (pinned host in, pinned host out, depth 2)",
                    "megapixels_per_s": ips_e2e * H * W / 1e6},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": tf / peaks["tf_sustained"], "traffic": traffic_bytes, "traffic_note": traffic_note,
                         "kernel": "implicit-GEMM conv family (all conv/deconv/linear/GDN launches)",
                         "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                         "share_of_step": conv_ms / total_ms,
                         "families_note": "eager per-launch profile of ONE sub-batch program (micro_batches of them run per step)",
                         "families_ms_per_step": {k: {"ms": v[0], "launches": v[1]} for k, v in fam.items()}},
        }
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            cips, _ = time_cpu_oracle(sd, steps=1, warmup=1, batch=2)
            line["cpu_baseline"] = {"value": cips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "1 step x 2 images 512x768 after a 1-image warm-up (oracle, torch CPU fp32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def profile_families(model, x_dev, dump=None):
    """Device time per kernel family for ONE forward, CUDA events around every launch (eager)."""
    import ctypes as C

    from resdsic_b200 import _lib
    plan = next(iter(model._plans.values()))
    plan.x.copy_(x_dev)
    prog = plan.prog
    L = _lib.lib()
    arr = prog._array()
    stream = torch.cuda.current_stream().cuda_stream
    names = {_lib.OP_CONV: "conv", _lib.OP_ATTN: "attn", _lib.OP_EB: "eb", _lib.OP_GC: "gc", _lib.OP_COPY: "copy", _lib.OP_PATCH: "patch",
             _lib.OP_LN: "ln"}
    n = len(prog.ops)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    opsz = C.sizeof(_lib.Op)
    base = C.addressof(arr)
    for rep in range(2):  # first rep warms caches
        evs[0].record()
        for i in range(n):
            if prog.ops[i].kind in _lib.SYNC_OPS:
                evs[i + 1].record()
                continue
            one = C.cast(base + i * opsz, C.POINTER(_lib.Op))
            _lib.check(L.rdsic_run_program(one, 1, stream, None, None))
            evs[i + 1].record()
        torch.cuda.synchronize()
    fam, per_op = {}, []
    for i in range(n):
        if prog.ops[i].kind in _lib.SYNC_OPS:
            continue
        k = names[prog.ops[i].kind]
        t = evs[i].elapsed_time(evs[i + 1])
        ms, cnt = fam.get(k, (0.0, 0))
        fam[k] = (ms + t, cnt + 1)
        rec = {"i": i, "kind": k, "ms": round(t, 4)}
        if k == "conv":
            c = prog.ops[i].u.conv
            flop = 2.0 * c.B * c.OH * c.OW * c.Cout * c.KH * c.KW * c.Cin
            rec.update(M=c.B * c.OH * c.OW, N=c.Cout, K=c.KH * c.KW * c.Cin, k=f"{c.KH}x{c.KW}s{c.stride}",
                       epi=c.epilogue, tc=int(c.w_dtype), tflops=round(flop / (t * 1e-3) / 1e12, 2))
        per_op.append(rec)
    if dump:
        with open(dump, "w") as fh:
            json.dump(per_op, fh, indent=0)
    return fam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RESDSIC_PRECISION", "bf16"), choices=["fp32", "bf16"])
    # 24 images x 1536 latent pixels = 144 of the 256-row tiles of the slice-loop GEMMs: one full wave of the 148
    # SMs (batch 16 fills 96 of them, batch 32 needs a second, 30 %-full wave).  Sweep: DESIGN.md section 5.
    ap.add_argument("--batch", type=int, default=24, help="images per GPU per step")
    ap.add_argument("--micro-batches", default="auto", help="sub-batches run as concurrent graphs (auto | 1 | 2 | 4 ...)")
    ap.add_argument("--model", default="cnn", choices=["cnn", "stf"], help="cnn = the headline (BASELINE.json) workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch device times of one step to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
