"""Per-LAYER table of one step: the per-launch ncu metrics (profiles/r2_ncu_step_metrics.csv) joined, launch by launch,
with the layer shapes of the step's program (`bench.py --dump-ops`, same 202-launch order).

    python profiles/layer_table.py profiles/r2_ncu_step_metrics.csv gpurun_out/ops.json > profiles/r2_layer_table.md
"""
import collections
import json
import sys

import step_table as S


def main(csv_path, ops_path):
    L = S.load(csv_path)
    starts = [i for i, l in enumerate(L) if "patchify" in l["name"]]
    step = L[starts[0]:starts[1]]
    ops = json.loads(open(ops_path).read().strip())
    assert len(step) == len(ops), (len(step), len(ops))
    T, TP = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
    DR, DW = "dram__bytes_read.sum", "dram__bytes_write.sum"
    agg = collections.OrderedDict()
    for l, o in zip(step, ops):
        k = (S.short(l["name"]), o.get("shape", o["kind"]))
        a = agg.setdefault(k, collections.defaultdict(float))
        a["n"] += 1
        a["t"] += l[T]
        a["tp"] += l.get(TP, 0.0) * l[T]
        a["bytes"] += l.get(DR, 0.0) + l.get(DW, 0.0)
        a["flop"] += o.get("flop", 0.0)
        a["alg"] += o.get("bytes", 0.0)
    tot = sum(a["t"] for a in agg.values())
    print(f"# One step (bench.py default: batch 48 x 512 x 768, bf16) by layer class: {len(step)} launches, {tot / 1e3:.2f} ms serialised under ncu\n")
    print("ncu per-launch duration / tensor-pipe activity / DRAM bytes; TFLOP/s = algorithmic FLOP / ncu duration; "
          "`DRAM / alg` = measured DRAM bytes over the layer's algorithmic bytes (in + out + weights).\n")
    print("| kernel | layer | launches | us each | share | tensor % | TFLOP/s | DRAM GB/s | DRAM / alg |")
    print("|---|---|---|---|---|---|---|---|---|")
    for (kern, shape), a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        tf = a["flop"] / a["t"] / 1e6 if a["flop"] else 0.0
        ratio = f"{a['bytes'] / a['alg']:.2f}" if a["alg"] else ""
        print(f"| `{kern}` | {shape} | {int(a['n'])} | {a['t'] / a['n']:.1f} | {100 * a['t'] / tot:.1f} % | {a['tp'] / a['t']:.1f} | "
              f"{tf:.0f} | {a['bytes'] / a['t'] / 1e3:.0f} | {ratio} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
