#!/usr/bin/env bash
# Same index and bytes, fewer queries: how much of the scan kernel follows the (query, list) pairs rather than the bytes
for q in 1024 256; do
  for m in 0; do
  QK_GRAPH=0 QK_SCAN_DBG=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --q $q 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('Q=$q dbg=$m kernel_ms', round(l['roofline']['kernel_ms'],4), 'alg MB', round(l['roofline']['algorithmic_bytes_per_launch']/1e6,1), 'GB/s', round(l['roofline']['achieved'],0), l['config']['scan_stats'])"
  done
done
