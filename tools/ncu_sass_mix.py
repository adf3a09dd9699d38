"""Opcode mix / stall samples per kernel from an .ncu-rep's source page:  python tools/ncu_sass_mix.py rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k, data, hdr = None, collections.defaultdict(list), None
for r in rows:
    if r and r[0] == "Kernel Name":
        k = r[1] + "#" + str(sum(1 for x in data if x.startswith(r[1])))
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if k and len(r) > 6:
        data[k].append(r)
for k, rs in data.items():
    si, ii, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot, ts = sum(int(r[ii]) for r in rs), sum(int(r[wi]) for r in rs)
    print(k, "warp-instr", tot, "samples", ts, "sass lines", len(rs))
    ops, samp = collections.Counter(), collections.Counter()
    for r in rs:
        f = r[si].split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        ops[op] += int(r[ii])
        samp[op] += int(r[wi])
    print("   " + "  ".join(f"{op} {100 * c / tot:.1f}%/{100 * samp[op] / max(ts, 1):.1f}%" for op, c in ops.most_common(top)))
