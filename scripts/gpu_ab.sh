#!/usr/bin/env bash
# A/B on one GPU box: each argument is TAG[:ENV=VAL[,ENV=VAL...]]; runs the quick bench per variant and prints a
# one-line summary (step ms, scan kernel ms, roofline frac, appended per query, parity). A variant tagged ...+ncu also
# gets an ncu launch list.
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%:*}"; envs=""; [ "$spec" != "$tag" ] && envs="${spec#*:}"
  want_ncu=0; case "$tag" in *+ncu) want_ncu=1; tag="${tag%+ncu}";; esac
  envline=$(echo "$envs" | tr ',' ' ')
  env $envline timeout 300 python bench.py --quick --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" "$envs" <<'PY'
import json, sys
tag, envs = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]; p = d["config"].get("parity", d.get("parity", {}))
    print(f"AB {tag:14s} step {d['ms_per_step']*1e3:7.1f} us  e2e {d['e2e']['ms_per_step']*1e3:7.1f} us  scan {r['kernel_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  "
          f"appended {d['scan_stats']['mean_appended_per_query']:.0f} rescanned {d['scan_stats']['queries_rescanned']} refresh {d['scan_stats'].get('threshold_refreshes')}/{d['scan_stats'].get('refresh_requests_dropped')}  "
          f"launches {d['gpu_launches_per_step']}  parity {d['parity'].get('ids_equal_oracle_port_16q')}/{d['parity'].get('dist_bit_equal_oracle_port_16q')}  [{envs}]")
except Exception as e:
    print(f"AB {tag}: FAILED {e!r}")
    print(open(f"gpurun_out/ab_{tag}.err").read()[-1500:])
PY
  if [ $want_ncu = 1 ]; then
    env $envline QK_BENCH_CUPROF=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_launch_$tag.log 2>&1
    python - "$tag" <<'PY'
import csv, sys
tag = sys.argv[1]
r = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
h = [i for i, x in enumerate(r) if x and x[0] == 'ID'][0]
hdr = r[h]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
rows = r[h + 1:]
half = rows[len(rows) // 2:]
tot = 0.0
for x in half:
    ns = float(x[iv].replace(',', '')); tot += ns
    print(f"   {x[ik][:58]:58s} {ns/1e3:8.1f} us")
print(f"   sum {tot/1e3:.1f} us")
PY
  fi
done
