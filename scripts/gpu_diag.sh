#!/usr/bin/env bash
# Diagnostics visit: full ncu capture of the step's non-scan kernels (source-level), launch list of one k-means assign
# call, warm timing of the k-means update kernels.
set -u
mkdir -p gpurun_out
TAG="${1:-diag}"
QK_BENCH_CUPROF=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"dense_refine|merge_refine|seed_thresholds|prefix_segments|scatter_pairs" -c 5 -f -o gpurun_out/prof_rest_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_rest_$TAG.log 2>&1
tail -2 gpurun_out/ncu_rest_$TAG.log
QK_PROBE_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/launches_assign_$TAG.csv python scripts/assign_speed.py 1000000 4096 > gpurun_out/assign_$TAG.log 2>&1
tail -2 gpurun_out/assign_$TAG.log
timeout 300 python scripts/update_speed.py 2>&1 | tail -6 | tee gpurun_out/update_$TAG.log
