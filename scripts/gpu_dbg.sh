#!/usr/bin/env bash
# Stage-elimination / variant timing of the scan kernel with the -DQK_STAGE_DEBUG build (quake_b200/lib/libquake_b200_dbg.so).
# QK_SCAN_DBG bits: 1 no selection, 2 no MMAs, 4 no a_lo split, 8 a_hi MMAs issued on TMA arrival, 16 npad always 32
for m in ${@:-0 8 16 24}; do
  QK_LIB_PATH=$PWD/quake_b200/lib/libquake_b200_dbg.so QK_SCAN_DBG=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('dbg=$m kernel_ms', round(l['roofline']['kernel_ms'],4), 'frac', round(l['roofline']['frac'],3), 'step_ms', round(l['ms_per_step'],3), 'e2e_ms', round(l['e2e']['ms_per_step'],3), 'appended', l['config']['scan_stats']['mean_appended_per_query'])"
done
