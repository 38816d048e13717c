#!/usr/bin/env bash
# Stage-elimination timing of the scan kernel (QK_SCAN_DBG bits: 1 no selection, 2 no MMAs, 4 no a_lo split)
for m in 0 1 2 3 4 6 7; do
  QK_SCAN_DBG=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('dbg=$m kernel_ms', round(l['roofline']['kernel_ms'],4), 'frac', round(l['roofline']['frac'],3), 'step_ms', round(l['ms_per_step'],3))"
done
