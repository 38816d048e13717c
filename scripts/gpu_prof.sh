#!/usr/bin/env bash
# ncu launch list of the timed steps + one full capture of the scan kernel (partition scan = 2nd scan launch of a step)
set -u
mkdir -p gpurun_out
TAG="${1:-x}"
QK_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
QK_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"scan_kernel|scan_mma|merge_refine" -c 4 -f -o gpurun_out/prof_scan_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
