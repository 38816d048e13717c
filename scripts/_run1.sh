#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
python scripts/assign_speed.py 1000000 4096 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d.get('build')); print(d['ms_per_step'], d['config'].get('build_s'), d.get('build_s'))"
