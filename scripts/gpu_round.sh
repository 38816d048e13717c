#!/usr/bin/env bash
# One GPU-box visit for the committed evidence of a round: parity tests, bench (both arms), ncu launch list + one full
# capture of the scan kernel (both launches of a step), the other BASELINE configs. Results land in gpurun_out/;
# scripts/collect_profiles.sh copies the summaries into profiles/.
set -u
mkdir -p gpurun_out
TAG="${1:-r02}"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_$TAG.txt
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_$TAG.err | tail -1 > gpurun_out/bench_$TAG.json
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value',d['value'],'step_ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'share',d['roofline']['kernel_share_of_step'],'parity',d['parity'].get('ids_equal_reference_1024q'),'cpu',d['cpu_baseline']['value'])"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref_$TAG.err | tail -1 | tee gpurun_out/bench_ref_$TAG.json | cut -c1-300
QK_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_launch_$TAG.log 2>&1
QK_BENCH_CUPROF=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"scan_mma" -c 2 -f -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
timeout 900 python scripts/run_configs.py c1 c3 c5 --tag $TAG 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
