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
    """`torch.rand(B,3,H,W)` from a private CPU generator (the synthetic input of the BASELINE configs)."""
    from resdsic_b200.utils import synthetic  # data generators only, no compute of the path
    return synthetic.rand_image(batch, H, W, seed=seed)


def make_weights(profile="refinit"):
    """Default "refinit": the reference constructor's own random init under torch.manual_seed(0) (BASELINE.json:
    "random-init weights"; bit-equal to the reference's, tests/test_refinit.py) -- the weights on which the
    bf16 tolerances of the north star are asserted un-relaxed.  "stress" / "lowrate": the hash-seeded profiles."""
    from resdsic_b200.utils import synthetic
    if profile == "refinit":
        return synthetic.refinit_state_dict(0)
    return synthetic.make_state_dict(seed=0, profile=profile)


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


CPU_SAMPLE_BATCH = 2  # images per CPU step: a bounded sample of the GPU arm's batch (0.4-0.5 s per image on 16 threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_weights(args.weights)
    b = CPU_SAMPLE_BATCH
    ips, per_step = time_cpu_oracle(sd, args.steps, args.warmup, batch=b)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cnn (WACNN N=192 M=320) forward 512x768, eval mode; bounded sample of the GPU arm's "
                               f"batch-{args.batch} step: {b} images per CPU step (images are independent, the CPU "
                               "rate does not depend on the batch)", "precision": "fp32", "weights": args.weights,
                   "batch_per_step": b, "image": [H, W]},
        "megapixels_per_s": ips * H * W / 1e6,
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {b} images 512x768, oracle (torch CPU fp32 restatement of "
                                   "the reference forward; the Python reference itself cannot travel to the GPU box)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_gpu_eager(sd, x_dev, steps=3, warmup=2):
    """The incumbent on the same B200: the SAME PyTorch graph the reference
    executes (the oracle's functional restatement: F.conv2d / conv_transpose2d / linear / matmul / softmax / erfc ...)
    in eager mode on the GPU, i.e. cuDNN + cuBLAS + ATen elementwise kernels -- fp32, TF32, and
    autocast(bf16) + channels_last.  A reported baseline (no kernel of this repo runs here)."""
    from oracle import wacnn_oracle as O
    dev = x_dev.device
    res = {}
    torch.backends.cudnn.benchmark = True
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    sd_cl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd_dev.items()}
    variants = (("fp32", False, None, sd_dev, x_dev),
                ("tf32", True, None, sd_dev, x_dev),
                ("bf16_autocast_channels_last", True, torch.bfloat16, sd_cl, x_dev.contiguous(memory_format=torch.channels_last)))
    for name, tf32, amp, w, x in variants:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            def step():
                if amp is None:
                    return O.forward(w, x)
                with torch.autocast("cuda", dtype=amp):
                    return O.forward(w, x)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            res[name] = {"images_per_s": x.shape[0] / (ms / 1e3), "ms_per_step": ms}
        except Exception as e:  # (out of memory on a small card, ...): report, do not fail the bench line
            res[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True
    return res


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
        try:  # one slice of the host cores per rank: the copy-issuing threads of the N processes do not migrate over each other
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
        except (AttributeError, OSError):
            pass

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    _lib.lib()  # fail loudly if the CUDA library is missing
    sd = make_weights(args.weights)
    if args.model == "stf":  # builder-defined model (no reference implementation): informational runs only
        from resdsic_b200.utils import synthetic as _w
        model = resdsic_b200.models["stf"]().eval()
        sd = _w.synth_state_dict(model.state_dict())
    else:
        model = resdsic_b200.WACNN().eval()
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).set_precision(args.precision)
    model.static_outputs = True  # results are consumed in place (ForwardPipeline snapshots them): no per-call clones
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

    # ---- the same with 8-bit images on the host side (ForwardPipeline(compact=True)): uint8 in, uint8 x_hat + the
    #      per-image rate out; the conversions and the rate reduction are kernels of this library (csrc/image_io.cu)
    u8_batches = [(hb * 255.0).round().to(torch.uint8).pin_memory() for hb in host_batches]
    pipe_c = ForwardPipeline(model, u8_batches[0], depth=2, compact=True)
    pipe_c.run([u8_batches[i % 2] for i in range(3)])
    barrier()
    e0.record()
    pipe_c.run([u8_batches[i % 2] for i in range(args.steps)])
    e1.record()
    barrier()
    ms_e2e_c = e0.elapsed_time(e1)

    from resdsic_b200.utils import max_over_ranks
    ms, ms_e2e, ms_e2e_c = max_over_ranks([ms, ms_e2e, ms_e2e_c], device=dev)

    # ---- per-kernel-family device time of one step (eager, CUDA events around every launch)
    fam = None
    if rank == 0:
        fam = profile_families(model, x_dev, args.dump_ops)

    if rank == 0:
        peaks = load_peaks()
        n_img = B * world * args.steps
        ips = n_img / (ms / 1e3)
        ips_e2e = n_img / (ms_e2e / 1e3)
        fam, per_op = fam
        conv_ms, conv_n = fam["conv"]
        total_ms = sum(v[0] for v in fam.values())
        flop_per_image = FLOP_PER_IMAGE
        if args.model != "cnn":  # no survey figure: sum 2*M*N*K over the program's GEMM descriptors
            plan = model._last_plan
            flop_per_image = sum(2.0 * o.u.conv.B * o.u.conv.OH * o.u.conv.OW * o.u.conv.Cout * o.u.conv.KH * o.u.conv.KW *
                                 o.u.conv.Cin for o in plan.prog.ops if o.kind == _lib.OP_CONV) / plan.sub_batch
        # the profile covers ONE sub-batch program (the model runs `micro_batches` of them concurrently)
        b_prog = model._last_plan.sub_batch
        fam_tf = b_prog * flop_per_image / (conv_ms / 1e3) / 1e12  # conv family: algorithmic FLOP of a step / its device time
        # dominant kernel = the (kernel instance, layer shape) class with the largest share of the step
        classes = {}
        for r in per_op:
            if r["kind"] != "conv":
                continue
            c = classes.setdefault(r["kernel"] + " " + r["shape"], {"ms": 0.0, "n": 0, "flop": r["flop"], "kernel": r["kernel"],
                                                                  "shape": r["shape"], "bytes": r["bytes"]})
            c["ms"] += r["ms"]
            c["n"] += 1
        dom = max(classes.values(), key=lambda c: c["ms"])
        dom_tf = dom["flop"] / (dom["ms"] / dom["n"] / 1e3) / 1e12
        step_tf = B * flop_per_image / (ms / args.steps / 1e3) / 1e12  # whole step (graph replay), all kernels
        traffic_bytes, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
        if os.path.exists(tpath):  # DRAM bytes of one launch of the dominant kernel, from this round's ncu --set full capture
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("kernel_key") == dom["kernel"] + " " + dom["shape"]:
                traffic_bytes = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_note = tj["source"]
        line = {
            "metric": METRIC if args.model == "cnn" else METRIC.replace("WACNN (-m cnn)", "STF (-m stf, builder-defined)"), "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": (f"cnn (WACNN N=192 M=320) forward 512x768, batch {B} per GPU, eval mode" if args.model == "cnn"
                                    else f"stf (builder-defined, N=192 M=384) forward 512x768, batch {B} per GPU, eval mode"),
                       "precision": args.precision, "weights": args.weights, "batch_per_gpu": B, "image": [H, W], "parallelism": f"dp{world}",
                       "l2": f"per-step activation working set (~{0.19 * B:.1f} GB at batch {B}) exceeds the 126 MB L2; no explicit flush",
                       "cuda_graph": bool(model.use_cuda_graph),
                       "micro_batches": len(model._last_plan.subs)},
            "megapixels_per_s": ips * H * W / 1e6,
            "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "api": "resdsic_b200.utils.ForwardPipeline.run (pinned host in, pinned host out, depth 2)",
                    "megapixels_per_s": ips_e2e * H * W / 1e6},
            "e2e_compact": {"value": n_img / (ms_e2e_c / 1e3), "unit": "images/s", "h2d_bytes_per_step": pipe_c.h2d_bytes,
                            "d2h_bytes_per_step": pipe_c.d2h_bytes,
                            "api": "ForwardPipeline(compact=True).run: uint8 images in; uint8 x_hat + per-image bits (fp64) out; "
                                   "conversions and rate reduction on the device"},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": dom_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": dom_tf / peaks["tf_sustained"], "frac_of_burst_peak": dom_tf / peaks["tf_burst"],
                         "peak_burst": peaks["tf_burst"],
                         "traffic": traffic_bytes, "traffic_note": traffic_note,
                         "kernel": dom["kernel"], "layer": dom["shape"], "launches_per_step": dom["n"],
                         "avg_launch_ms": dom["ms"] / dom["n"], "algorithmic_flop_per_launch": dom["flop"],
                         "algorithmic_bytes_per_launch": dom["bytes"], "share_of_step": dom["ms"] / total_ms,
                         "peak_source": peaks["source"] + ": sustained bf16 (kernel timed inside a long step) and burst",
                         "timing": "CUDA events around every launch of one eager pass of the step's program (second of two passes)",
                         "conv_family": {"achieved": fam_tf, "frac": fam_tf / peaks["tf_sustained"],
                                         "frac_of_burst_peak": fam_tf / peaks["tf_burst"], "share_of_step": conv_ms / total_ms,
                                         "launches_per_step": conv_n},
                         "whole_step": {"achieved": step_tf, "frac": step_tf / peaks["tf_sustained"],
                                        "frac_of_burst_peak": step_tf / peaks["tf_burst"],
                                        "note": "B x 413.22 GFLOP / graph-replay step time (every kernel, launch gaps included)"},
                         "families_ms_per_step": {k: {"ms": v[0], "launches": v[1]} for k, v in fam.items()}},
        }
        if world == 1 and not args.no_eager_baseline and args.model == "cnn":
            eager = time_gpu_eager(sd, x_dev)
            best = max((v["images_per_s"] for v in eager.values() if "images_per_s" in v), default=None)
            line["gpu_eager_baseline"] = {
                "what": "the same PyTorch graph (oracle restatement of the reference forward) in eager mode on this GPU: "
                        "cuDNN / cuBLAS / ATen kernels, same batch and image size, inputs resident, CUDA-event timed",
                "batch": B, "variants": eager, "best_images_per_s": best,
                "speedup_over_best": (ips / best) if best else None}
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            cips, _ = time_cpu_oracle(sd, steps=2, warmup=1, batch=CPU_SAMPLE_BATCH)
            line["cpu_baseline"] = {"value": cips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"2 steps x {CPU_SAMPLE_BATCH} images 512x768 after a warm-up step (oracle, torch CPU fp32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- training workload (BASELINE config 4)
TRAIN_METRIC = "WACNN (-m cnn) rate-distortion training step images/s at 256x256 (lambda = 0.0035, fp32)"
TRAIN_H = TRAIN_W = 256


def _train_objects(dev, world):
    """Model, criterion and the two optimisers exactly as train.py:55-89 / training/step.py:18-56 set them up:
    Adam over every parameter but the `.quantiles`, a second Adam over the `.quantiles`, clip 1.0."""
    import resdsic_b200
    from resdsic_b200.training import GradBucketReducer, RateDistortionLoss
    net = resdsic_b200.WACNN().train()
    net.load_state_dict(make_weights("refinit"), strict=True)
    net = net.to(dev)
    main = [p for n, p in net.named_parameters() if not n.endswith(".quantiles") and p.requires_grad]
    aux = [p for n, p in net.named_parameters() if n.endswith(".quantiles") and p.requires_grad]
    opt = torch.optim.Adam(main, lr=1e-4)
    aux_opt = torch.optim.Adam(aux, lr=1e-3)
    crit = RateDistortionLoss(lmbda=0.0035)
    reducer = GradBucketReducer(main + aux) if world > 1 else None
    return net, crit, opt, aux_opt, reducer, main


def _train_step(net, crit, opt, aux_opt, reducer, main, x, ev=None):
    """training/step.py:36-56, one batch.  `ev`: optional dict of CUDA events recorded at the phase boundaries."""
    def mark(k):
        if ev is not None:
            ev[k].record()
    opt.zero_grad(set_to_none=True)
    aux_opt.zero_grad(set_to_none=True)
    mark("t0")
    out = net(x)
    oc = crit(out, x)
    mark("fwd")
    oc["loss"].backward()
    aux = net.aux_loss()
    aux.backward()
    mark("bwd")
    if reducer is not None:
        reducer.finish()  # waits for the bucketed all-reduces that overlapped the backward pass
    mark("red")
    aux_opt.step()
    torch.nn.utils.clip_grad_norm_(main, 1.0)
    opt.step()
    mark("opt")
    return oc["loss"]


def time_gpu_eager_train(x_dev, steps=2, warmup=1):
    """The incumbent for config 4 on this GPU: the same PyTorch graph (oracle restatement of the training-mode forward)
    with torch autograd + cuDNN/cuBLAS backward, RD loss, Adam -- fp32 and TF32.  (The oracle's round() passes no
    straight-through gradient, so its backward graph is, if anything, slightly smaller than the reference's.)"""
    import math
    from oracle import wacnn_oracle as O
    dev = x_dev.device
    res = {}
    B = x_dev.shape[0]
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        try:
            sd = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in make_weights("refinit").items()}
            params = [v for v in sd.values() if v.requires_grad]
            opt = torch.optim.Adam(params, lr=1e-4)
            noise = {"y": torch.rand(B, 320, TRAIN_H // 16, TRAIN_W // 16, device=dev) - 0.5,
                     "z": torch.rand(B, 192, TRAIN_H // 64, TRAIN_W // 64, device=dev) - 0.5}

            def step():
                opt.zero_grad(set_to_none=True)
                out = O.forward.__wrapped__(sd, x_dev, noise=noise)  # (the oracle's forward is decorated no_grad)
                npx = B * TRAIN_H * TRAIN_W
                bpp = sum(torch.log(l).sum() / (-math.log(2) * npx) for l in out["likelihoods"].values())
                loss = 0.0035 * 255 ** 2 * torch.nn.functional.mse_loss(out["x_hat"], x_dev) + bpp
                loss.backward()
                torch.nn.utils.clip_grad_norm_(params, 1.0)
                opt.step()
            for _ in range(warmup):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            res[name] = {"images_per_s": B / (ms / 1e3), "ms_per_step": ms}
        except Exception as e:
            res[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True
    return res


def run_train(args):
    """`--workload train`: BASELINE config 4 -- one RD training step (forward, loss, backward, aux step, clip, Adam) per
    GPU on a batch of 16 x 3 x 256 x 256, fp32 kernels of this library; N > 1: one process per GPU, the bucketed
    gradient all-reduce (NCCL) overlapping the backward pass is the only collective.  An extra line, not the headline."""
    import torch.distributed as dist

    from resdsic_b200 import _lib
    from resdsic_b200.training import functions as Fn
    from resdsic_b200.utils import max_over_ranks, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    _lib.lib()
    B = args.batch
    net, crit, opt, aux_opt, reducer, main = _train_objects(dev, world)
    x_host = synthetic.rand_image(B, TRAIN_H, TRAIN_W, seed=300 + rank).pin_memory()
    x_dev = x_host.to(dev)
    for _ in range(max(3, args.warmup)):
        _train_step(net, crit, opt, aux_opt, reducer, main, x_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    keys = ("t0", "fwd", "bwd", "red", "opt")
    evs = [{k: torch.cuda.Event(enable_timing=True) for k in keys} for _ in range(args.steps)]
    n0 = Fn.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        _train_step(net, crit, opt, aux_opt, reducer, main, x_dev, evs[i])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = Fn.LAUNCHES[0] - n0
    clocks = sampler.stop() if rank == 0 else None
    phase = {b: sum(e[a].elapsed_time(e[b]) for e in evs) / args.steps for a, b in zip(keys[:-1], keys[1:])}

    # end to end: the batch comes from pinned host memory every step and the loss is read back on the host
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    barrier()
    e0.record()
    for i in range(args.steps):
        xb = x_host.to(dev, non_blocking=True)
        loss = _train_step(net, crit, opt, aux_opt, reducer, main, xb)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        float(loss_host)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # the collective alone (not overlapped): all-reduce of every gradient bucket, back to back
    ar_ms, grad_bytes = None, sum(p.numel() * 4 for p in net.parameters())
    if world > 1:
        flats = [torch.zeros(sum(p.numel() for p in b), device=dev) for b in reducer.buckets]
        for _ in range(2):
            for f in flats:
                dist.all_reduce(f)
        barrier()
        e0.record()
        for f in flats:
            dist.all_reduce(f)
        e1.record()
        barrier()
        ar_ms = e0.elapsed_time(e1)
        del flats
    vals = [ms, ms_e2e, phase["fwd"], phase["bwd"], phase["red"], phase["opt"]] + ([ar_ms] if ar_ms is not None else [])
    vals = max_over_ranks(vals, device=dev)
    if rank == 0:
        ms, ms_e2e = vals[0], vals[1]
        n_img = B * world * args.steps
        line = {
            "metric": TRAIN_METRIC, "value": n_img / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cnn (WACNN N=192 M=320) RD training step (BASELINE config 4): batch {B} x 3 x 256 x 256 per "
                                   "GPU, lambda 0.0035, forward + loss + backward + aux step + clip 1.0 + Adam",
                       "precision": "fp32", "weights": "refinit", "batch_per_gpu": B, "image": [TRAIN_H, TRAIN_W],
                       "parallelism": f"dp{world}", "collective": "bucketed NCCL all-reduce of the gradients (25 MB buckets) on a side "
                                                                  "stream, overlapping backward" if world > 1 else "none (1 GPU)",
                       "l2": "activations + saved tensors of a step exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": n_img / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "api": "model(x) / RateDistortionLoss / loss.backward() / optimizer.step() "
                                                    "(training/step.py's calls), batch from pinned host memory, loss read on the host"},
            "gpu_launches": launches, "launches_per_step": launches // args.steps, "clocks": clocks,
            "phases_ms": {"forward_and_loss": vals[2], "backward": vals[3], "allreduce_exposed_after_backward": vals[4],
                          "aux_step_clip_adam": vals[5],
                          "allreduce_alone": vals[6] if ar_ms is not None else None, "gradient_bytes": grad_bytes},
        }
        if world == 1 and not args.no_eager_baseline:
            eager = time_gpu_eager_train(x_dev)
            best = max((v["images_per_s"] for v in eager.values() if "images_per_s" in v), default=None)
            line["gpu_eager_baseline"] = {"what": "the same training step as torch eager autograd on this GPU (oracle graph, cuDNN / "
                                                  "cuBLAS), same batch", "variants": eager, "best_images_per_s": best,
                                          "speedup_over_best": (line["value"] / best) if best else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def profile_families(model, x_dev, dump=None):
    """Device time per kernel family for ONE forward, CUDA events around every launch (eager)."""
    import ctypes as C

    from resdsic_b200 import _lib
    plan = model._last_plan
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
            M = c.B * c.OH * c.OW
            flop = 2.0 * M * c.Cout * c.KH * c.KW * c.Cin
            n_out = c.Cout
            if c.tail_mode:  # fused second GEMM (GDN / IGDN / ResidualUnit tail): K2 = Cout, N2 = tail_n
                flop += 2.0 * M * c.Cout * c.tail_n
                n_out = c.tail_n
            esz = lambda v: 0 if not v.ptr else (2 if v.dtype == _lib.BF16 else 4)
            # algorithmic bytes: input once, outputs once, residual / gate operands once, weights once
            nbytes = (c.B * c.H * c.W * c.Cin * esz(c.in_) + M * n_out * (esz(c.out) + esz(c.out2) + esz(c.out3) +
                      esz(c.res) + esz(c.aux)) + c.Cout * c.KH * c.KW * c.Cin * 2 + c.tail_n * c.Cout * 2)
            ru_pair = (c.tail_mode == 3 and c.tail_n == 2 * c.Cout and 5 * c.Cout <= 512 and 64 < c.Cin <= 128 and c.KH * c.KW >= 2
                       and c.stride == 1 and os.environ.get("RDSIC_RU_PAIR", "1") != "0")  # ru_pair_bf16.cu's eligibility rule
            kern = ("ru_pair_tc_kernel" if ru_pair else f"conv_gdn_tc_kernel<{c.tail_mode}>" if c.tail_mode else
                    (f"conv_tc_kernel<EPI={c.epilogue}>" if c.w_dtype == _lib.BF16 else "conv_f32_kernel"))
            shape = f"M={M} N={c.Cout} K={c.KH * c.KW * c.Cin} {c.KH}x{c.KW}s{c.stride}" + (f" tailN={c.tail_n}" if c.tail_mode else "")
            rec.update(M=M, N=c.Cout, K=c.KH * c.KW * c.Cin, k=f"{c.KH}x{c.KW}s{c.stride}", kernel=kern, shape=shape, flop=flop,
                       bytes=nbytes, epi=c.epilogue, tc=int(c.w_dtype), tflops=round(flop / (t * 1e-3) / 1e12, 2))
        per_op.append(rec)
    if dump:
        with open(dump, "w") as fh:
            json.dump(per_op, fh, indent=0)
    return fam, per_op


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="forward", choices=["forward", "train"],
                    help="forward = the headline (BASELINE configs[2]); train = the RD training step of config 4 (extra line)")
    ap.add_argument("--precision", default=os.environ.get("RESDSIC_PRECISION", "bf16"), choices=["fp32", "bf16"])
    # forward: 48 images x 1536 latent pixels = 288 of the 256-row tiles of the slice-loop GEMMs = two full waves of the
    # 74 CTA pairs (24 -> one wave: 2 % slower per image, the per-launch prologues weigh twice as much; 64 / 96 -> equal
    # to 48 within noise: the step runs into the board's power cap).  Sweep: DESIGN.md section 5.  train: 16 (config 4).
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: 48 forward, 16 train)")
    ap.add_argument("--micro-batches", default="auto", help="sub-batches run as concurrent graphs (auto | 1 | 2 | 4 ...)")
    ap.add_argument("--model", default="cnn", choices=["cnn", "stf"], help="cnn = the headline (BASELINE.json) workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the eager-PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--weights", default="refinit", choices=["refinit", "stress", "lowrate"],
                    help="refinit = the reference constructor's random init (seed 0); stress / lowrate = hash-seeded profiles")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch device times of one step to this JSON file")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 16 if args.workload == "train" else 48
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
