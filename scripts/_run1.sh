#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_index.py tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_ab.sh carve+ncu nocarve:QK_CARVEOUT=0 carve2 nocarve2:QK_CARVEOUT=0
QK_BENCH_CUPROF=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"dense_refine|merge_refine" -c 2 -f -o gpurun_out/prof_rest_d2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_rest_d2.log 2>&1
tail -2 gpurun_out/ncu_rest_d2.log
