"""Summarise an `ncu --page source --csv` dump: hottest SASS lines by stall samples with the dominant
stall reason.  usage: python tests/ncu_hot.py file.csv [N]"""
import csv
import sys


def main(path, n=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, body = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    si = hdr.index("# Samples")
    src = hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(float(r[si] or 0) for r in body)
    print("total samples", tot, "instructions", len(body))
    order = sorted(range(len(body)), key=lambda k: -float(body[k][si] or 0))[:n]
    for k in order:
        r = body[k]
        st = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
        print(f"{float(r[si]):8.0f} {100 * float(r[si]) / tot:5.1f}%  #{k:5d} {r[src].strip()[:90]:90s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
