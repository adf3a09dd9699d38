"""DRAM traffic per launch of a kernel family from an `ncu --set full` capture (read here, no GPU):

    python tools/traffic_json.py capture.ncu-rep <family name in bench.py's kernel table> "<how it was captured>" [kernel-name regex] [out.json]

Merges {family: {traffic_bytes_per_launch, launches_captured, per_launch: [{us, read, write}], how}} into
profiles/r02_traffic.json (default; bench.py reads r02 first, then r01), which bench.py reads for `roofline.traffic`.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}[unit]


def main():
    import re
    rep, family, how = sys.argv[1], sys.argv[2], sys.argv[3]
    pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] else None
    out_name = sys.argv[5] if len(sys.argv) > 5 else "r02_traffic.json"
    # a .ncu-rep, or its `ncu -i rep --page raw --csv --print-units base` export
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    per = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pat is not None and not pat.search(d["Kernel Name"]):
            continue
        rec = {"kernel": d["Kernel Name"][:90], "us": round(to_us(d["gpu__time_duration.sum"], units["gpu__time_duration.sum"]), 3),
               "read": to_bytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"]),
               "write": to_bytes(d["dram__bytes_write.sum"], units["dram__bytes_write.sum"])}
        # what THIS launch asked of the L2 (sectors x 32 B): the yardstick for its own DRAM traffic -- a family's launches differ in size,
        # and the captured ones need not be average
        rk, wk = "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"
        if d.get(rk):
            rec["l2_read_requested"] = 32 * float(d[rk].replace(",", ""))
        if d.get(wk):
            rec["l2_write_requested"] = 32 * float(d[wk].replace(",", ""))
        per.append(rec)
    path = os.path.join(ROOT, "profiles", out_name)
    data = json.load(open(path)) if os.path.exists(path) else {}
    req = [p.get("l2_read_requested", 0) + p.get("l2_write_requested", 0) for p in per]
    data[family] = {"traffic_bytes_per_launch": round(sum(p["read"] + p["write"] for p in per) / max(1, len(per))),
                    "l2_requested_bytes_per_launch": round(sum(req) / max(1, len(per))) if any(req) else None,
                    "launches_captured": len(per), "per_launch": per, "how": how}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data[family], indent=1))


if __name__ == "__main__":
    main()
