#!/usr/bin/env bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sl in 256 512 1024 2048; do
  QK_SEG_LEN=$sl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('seg_len=$sl step_ms', round(l['ms_per_step'],4), 'e2e_ms', round(l['e2e']['ms_per_step'],4), 'kernel_ms', round(l['roofline']['kernel_ms'],4), l['config']['parity'])"
done
