#!/usr/bin/env bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list + one full capture of the scan kernel.
set -u
mkdir -p gpurun_out
TAG="${1:-r1}"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.txt
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_$TAG.err | tail -1 | tee gpurun_out/bench_$TAG.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref_$TAG.err | tail -1 | tee gpurun_out/bench_ref_$TAG.json
QK_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
QK_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"scan_mma" -c 2 -f -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
