#!/bin/bash
# launch lists of a tgnet_fps step: usage gpu_tgn_list.sh [nograd]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MODE=${1:-grad}
OUT=gpurun_out/r2_launches_tgnet_${MODE}.csv
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "measured_step/" --csv --log-file $OUT python scripts/ncu_tgn_driver.py b200 $MODE > gpurun_out/ncu_tgn.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_tgn.log; wc -l $OUT
