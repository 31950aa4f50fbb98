"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the Blackwell-only opcodes that prove the
hot path is tcgen05 / TMEM / TMA code.  usage: python profiles/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "resdsic_b200", "lib", "libresdsic_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "MUFU", "LDG", "STG",
         "LDS", "STS", "LDGSTS", "ELECT", "UCGABAR", "FFMA")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur = collections.OrderedDict(), None
    names = iter(demangle)
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            kernels[cur][m.group(1).split(".")[0]] += 1
            kernels[cur]["_total"] += 1
    tot = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels (cuobjdump -sass, sm_100a)")
    print("# columns: total SASS instructions | " + " ".join(WATCH))
    for k, c in kernels.items():
        short = re.sub(r"\(anonymous namespace\)::", "", k)
        short = re.sub(r"\((int|bool)\)", "", short).replace("void ", "").replace("<unnamed>::", "")
        short = re.sub(r"\(.*", "", short)
        print(f"{short[:70]:70s} {c['_total']:6d} | " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
        tot.update(c)
    print(f"{'TOTAL':70s} {tot['_total']:6d} | " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))


if __name__ == "__main__":
    main()
