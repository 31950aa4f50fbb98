"""Summarise one `ncu --set full` capture (raw CSV page) into the small JSON bench.py reads for `roofline.traffic`.
usage: python profiles/make_traffic_json.py RAW.csv "<kernel key as bench.py prints it>" OUT.json"""
import csv
import json
import sys


def main(raw, key, out):
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def num(name, scale_units):
        u, v = col[name]
        return float(v.replace(",", "")) * scale_units.get(u, 1.0)
    byte_u = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    time_u = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    js = {
        "kernel_key": key,
        "kernel_name": col.get("Kernel Name", ("", ""))[1],
        "duration_us_under_ncu": num("gpu__time_duration.sum", time_u),
        "dram_bytes_read": num("dram__bytes_read.sum", byte_u),
        "dram_bytes_write": num("dram__bytes_write.sum", byte_u),
        "l2_to_sm_bytes": num("l1tex__m_xbar2l1tex_read_bytes.sum", byte_u),
        "tensor_pipe_active_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", {}),
        "lts_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed", {}),
        "registers_per_thread": num("launch__registers_per_thread", {}),
        "source": "ncu --set full --clock-control none, one launch inside `python bench.py --steps 2 --warmup 3` (" + raw + ")",
    }
    with open(out, "w") as fh:
        json.dump(js, fh, indent=1)
    print(json.dumps(js))


if __name__ == "__main__":
    main(*sys.argv[1:4])
