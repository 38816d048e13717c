for cfg in "1000000 4096" "500000 2048" "250000 1024" "2000000 8192"; do
  set -- $cfg
  for m in 0 7; do
  QK_GRAPH=0 QK_SCAN_DBG=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --n $1 --nlist $2 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('n=$1 nlist=$2 dbg=$m kernel_ms', round(l['roofline']['kernel_ms'],4), 'alg MB', round(l['roofline']['algorithmic_bytes_per_launch']/1e6,1), 'GB/s', round(l['roofline']['achieved'],0))"
  done
done
