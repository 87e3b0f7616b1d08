#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_mlp_tc8_kernel|sa_mlp_tc_kernel" -c 2 -o gpurun_out/prof_r2_tc8 -f python scripts/ncu_mlp_driver.py > gpurun_out/ncu_tc8.log 2>&1
tail -5 gpurun_out/ncu_tc8.log
ncu -i gpurun_out/prof_r2_tc8.ncu-rep --page raw --csv > gpurun_out/r2_ncu_tc8_raw.csv 2>/dev/null
ls -la gpurun_out/prof_r2_tc8.ncu-rep gpurun_out/r2_ncu_tc8_raw.csv
