"""Per-kernel-class table of ONE step from the per-launch ncu metric list written by profiles/collect_r2_step.sh.

    python profiles/step_table.py gpurun_out/r2_ncu_step_metrics.csv > profiles/r2_kernel_table.md

A step starts at the `patchify` launch (the first kernel of the forward program) and runs to the launch before the
next one.  ncu serialises the launches and replays each a few times with cold caches: durations are for SHARES, the
tensor-pipe / DRAM columns are per-kernel facts.  HBM peak = MEASURED_PEAKS.json (copy bandwidth)."""
import collections
import csv
import json
import os
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    hdr = rows[hi]
    idc, kn, mn, mu, mv = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        L = launches.setdefault(r[idc], {"name": r[kn]})
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[mu]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
                 "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}.get(unit, 1.0)
        L[r[mn]] = v * scale
    return list(launches.values())


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    return re.sub(r"\(.*$", "", name)


def main(path):
    L = load(path)
    starts = [i for i, l in enumerate(L) if "patchify" in l["name"]]
    if len(starts) < 2:
        raise SystemExit(f"need two step starts in the capture, found {len(starts)} (of {len(L)} launches)")
    step = L[starts[0]:starts[1]]
    here = os.path.dirname(os.path.abspath(__file__))
    peaks = json.load(open(os.path.join(here, "..", "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    T, TP, DR, DW, DP, IS, RG, XB = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                     "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                     "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                                     "l1tex__m_xbar2l1tex_read_bytes.sum")
    agg = collections.OrderedDict()
    for l in step:
        a = agg.setdefault(short(l["name"]), collections.defaultdict(float))
        t = l.get(T, 0.0)
        a["n"] += 1
        a["t"] += t
        a["tp_t"] += l.get(TP, 0.0) * t
        a["is_t"] += l.get(IS, 0.0) * t
        a["bytes"] += l.get(DR, 0.0) + l.get(DW, 0.0)
        a["xbar"] += l.get(XB, 0.0)
        a["regs"] = max(a["regs"], l.get(RG, 0.0))
    tot = sum(a["t"] for a in agg.values())
    print(f"# One step of `bench.py` (bench.py default: batch 48 x 512 x 768, bf16), per kernel class: {len(step)} launches, {tot / 1e3:.2f} ms serialised under ncu\n")
    print(f"HBM peak (measured copy bandwidth) = {hbm:.0f} GB/s.  `tensor %` = sm__pipe_tensor_cycles_active (time-weighted), `DRAM GB/s` = "
          "(dram read + write bytes) / duration, `L2->SM MB` = l1tex__m_xbar2l1tex_read_bytes per launch.\n")
    print("| kernel | launches | us total | share | tensor % | issue % | DRAM MB / launch | DRAM GB/s | of HBM peak | L2->SM MB / launch | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = a["bytes"] / a["t"] / 1e3 if a["t"] else 0.0
        print(f"| `{k}` | {int(a['n'])} | {a['t']:.0f} | {100 * a['t'] / tot:.1f} % | {a['tp_t'] / a['t']:.1f} | {a['is_t'] / a['t']:.1f} | "
              f"{a['bytes'] / a['n'] / 1e6:.1f} | {gbs:.0f} | {gbs / hbm:.2f} | {a['xbar'] / a['n'] / 1e6:.1f} | {int(a['regs'])} |")


if __name__ == "__main__":
    main(sys.argv[1])
