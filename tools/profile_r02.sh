#!/bin/bash
# Round-2 ncu evidence (run on the GPU box under gpurun, AFTER the same bench command exited 0 without ncu):
#   bash tools/profile_r02.sh [a|b|c|d ...]  -> gpurun_out/r02_*.csv / *.ncu-rep (kept under 64 MiB in total), summarised here into
#                                              profiles/ by tools/launch_summary.py, tools/ncu_summary.py and tools/traffic_json.py
# Cost note: `--set full` replays every kernel ~40 times, ~5-8 s per captured launch inside this program; 130 launches took 17 minutes.
set -u
mkdir -p gpurun_out
WHAT="${*:-a b c d}"
B="python bench.py --steps 2 --warmup 6 --batches 2 --no-extra --no-cpu-baseline --no-clock-sampler --profile-run"
$B > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || { echo "plain run failed"; exit 1; }
for w in $WHAT; do
  S=$(date +%s)
  case $w in
  a) # launch list of steady-state steps: time only, no replay
     ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5200 -c 2600 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_a.log 2>&1 ;;
  b) # the GEMM families inside the bench (one stretch of consecutive launches: forward NT with fused epilogues, backward-data NN, weight-gradient TN)
     ncu --set full --clock-control none -k regex:bf16_gemm_kernel --launch-skip 2100 -c 36 -f -o gpurun_out/r02_gemm $B > gpurun_out/r02_ncu_b.log 2>&1 ;;
  b2) # ... and a stretch of the backward pass: backward-data NN and weight-gradient TN launches alternate
     ncu --set full --clock-control none -k regex:bf16_gemm_kernel --launch-skip 2270 -c 14 -f -o gpurun_out/r02_gemm_bwd $B > gpurun_out/r02_ncu_b2.log 2>&1 ;;
  c) # window attention (tcgen05 tiles + warp kernels), forward and backward
     ncu --set full --clock-control none -k regex:attn_ --launch-skip 354 -c 12 -f -o gpurun_out/r02_attn $B > gpurun_out/r02_ncu_c.log 2>&1 ;;
  d) # the small kernels around the path: memory sections only
     ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
         -k 'regex:vox_|part_|chamfer|subm_|strided_|ln_bwd|colsum|bn2d|bn_|cast_multi|gather_rows|scatter_rows|seg_max|densify|rows_dense|onehot|pos_table' \
         --launch-skip 1500 -c 70 -f -o gpurun_out/r02_small $B > gpurun_out/r02_ncu_d.log 2>&1 ;;
  esac
  echo "$w rc=$? took $(( $(date +%s) - S )) s"
  # what travels back: the raw-page CSV export (base units) of every report; the report itself only when it is small
  for r in gpurun_out/*.ncu-rep; do
    [ -f "$r" ] || continue
    ncu -i "$r" --page raw --csv --print-units base > "${r%.ncu-rep}.csv" 2>/dev/null
    [ $(stat -c %s "$r") -gt 12000000 ] && rm -f "$r"
  done
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv; du -sh gpurun_out
