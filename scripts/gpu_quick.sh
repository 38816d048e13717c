#!/usr/bin/env bash
# Quick GPU iteration: parity tests, then the bench without the CPU baseline, then (optionally) an ncu launch list.
set -u
mkdir -p gpurun_out
TAG="${1:-q}"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_$TAG.err | tail -1 | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
if [ "${2:-}" = "ncu" ]; then
  QK_BENCH_CUPROF=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
  QK_BENCH_CUPROF=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"scan_mma" -c 2 -f -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_$TAG.log
fi
