#!/usr/bin/env bash
# Copy the summaries of one scripts/gpu_round.sh visit from gpurun_out/ (scratch) into profiles/ (tracked).
# usage: bash scripts/collect_profiles.sh <tag of the run> <name in profiles, e.g. r02>
set -u
TAG="$1"; OUT="${2:-r02}"
cp gpurun_out/bench_$TAG.json profiles/${OUT}_bench.json
cp gpurun_out/bench_ref_$TAG.json profiles/${OUT}_bench_reference.json
cp gpurun_out/pytest_gpu_$TAG.txt profiles/${OUT}_pytest_gpu.txt
cp gpurun_out/launches_$TAG.csv profiles/${OUT}_ncu_launches.csv
[ -f gpurun_out/configs_${TAG}_c1_c3_c5.json ] && cp gpurun_out/configs_${TAG}_c1_c3_c5.json profiles/${OUT}_configs_c1_c3_c5.json
python scripts/ncu_traffic.py gpurun_out/prof_scan_$TAG.ncu-rep profiles/${OUT}_ncu_full_scan_mma.json
ncu -i gpurun_out/prof_scan_$TAG.ncu-rep --page details 2>/dev/null | grep -v "^ *$" > profiles/${OUT}_ncu_details_scan_mma.txt
# SASS evidence of the tensor-core / TMA path in the shipped library
{
  echo "# cuobjdump -sass quake_b200/lib/libquake_b200.so (sm_100a), instruction counts per kernel"
  cuobjdump -sass quake_b200/lib/libquake_b200.so 2>/dev/null | awk '
    /Function : /{name=$3}
    /UTCHMMA/{mma[name]++} /UTMALDG/{tma[name]++} /UBLKCP/{blk[name]++} /LDTM/{ldtm[name]++} /STTM/{sttm[name]++} /UTCBAR/{bar[name]++} /SYNCS/{syncs[name]++} /REDUX/{redux[name]++}
    END{for(n in syncs) printf "%-90s UTCHMMA %3d  UTMALDG %2d  UBLKCP %2d  LDTM %2d  STTM %2d  UTCBAR %2d  SYNCS(mbarrier) %3d  REDUX %3d\n", substr(n,1,90), mma[n], tma[n], blk[n], ldtm[n], sttm[n], bar[n], syncs[n], redux[n]}' | sort
  echo
  cuobjdump --dump-resource-usage quake_b200/lib/libquake_b200.so 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/Fatbin.*Function //' | cut -c1-220
} > profiles/${OUT}_sass_summary.txt
ls -la profiles | grep "${OUT}_"
