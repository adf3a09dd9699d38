"""Summarise an .ncu-rep (read here, no GPU):  python tools/ncu_summary.py rep [kernel-substring]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "derived__smsp__sass_thread_inst_executed_op_local",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main():
    rep, sub = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    # a .ncu-rep, or the `ncu -i rep --page raw --csv` export of one (the export travels back from the GPU box when the report is too big)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if sub not in d["Kernel Name"]:
            continue
        print(f"{'Kernel Name':75s} {d['Kernel Name']}")
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k]} {units[hdr.index(k)]}")
        stalls = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(d[h]) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]}
        print("stalls (warps per issue): " + ", ".join(f"{k}={v:.2f}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
        print("-----")


if __name__ == "__main__":
    main()
