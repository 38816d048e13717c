#!/usr/bin/env bash
# per-role wait breakdown of the scan kernel (needs quake_b200/lib/libquake_b200_dbg.so, built with -DQK_STAGE_DEBUG)
QK_LIB_PATH=$PWD/quake_b200/lib/libquake_b200_dbg.so timeout 300 python scripts/role_probe.py "$@" 2>&1 | tail -20 | tee gpurun_out/roles.txt
