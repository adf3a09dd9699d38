"""Summarises the ncu launch list of ONE pretraining step (read here, no GPU):

    ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip S -c N --csv --log-file launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline          (on the GPU box, under gpurun)
    python tools/launch_summary.py launches.csv [title] > profiles/rNN_launches_step_summary.txt

A step is the span from the first `vox_mark_kernel` of a voxelisation pair (current + previous frame) to the first one of
the next pair.  Per kernel (template arguments kept, parameter list dropped): total time, share, launches, average.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("tmae::", "").replace("at::", "")
    depth, out = 0, []
    for ch in name:        # drop the parameter list: cut at the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out)[:110]


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else "one pretraining step"
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3))
    marks = [i for i, (n, _) in enumerate(rows) if n.startswith("vox_mark_kernel")]
    starts = [m for j, m in enumerate(marks) if j == 0 or m - marks[j - 1] > 100]   # first mark of each (cur, prev) pair
    if len(starts) < 2:
        raise SystemExit(f"need two step boundaries, found {len(starts)} in {len(rows)} launches")
    step = rows[starts[0]:starts[1]]
    total = sum(t for _, t in step)
    agg = defaultdict(lambda: [0.0, 0])
    for n, t in step:
        agg[n][0] += t
        agg[n][1] += 1
    print(f"{title}: {len(step)} launches, {total / 1e3:.2f} ms of kernel time (ncu --metrics gpu__time_duration.sum "
          f"--clock-control none: cold cache, serialised)")
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t:9.1f} us {100 * t / total:5.1f}% {c:5d}x {t / c:8.1f} avg  {n}")


if __name__ == "__main__":
    main()
