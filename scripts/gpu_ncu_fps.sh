#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fps_bucket_kernel" -s 1 -c 1 -o gpurun_out/prof_r2_fps_v1 -f python scripts/ncu_fps_driver.py 1 > gpurun_out/ncu_fps.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fps_bucket_kernel2" -s 1 -c 1 -o gpurun_out/prof_r2_fps_v2 -f python scripts/ncu_fps_driver.py 1 >> gpurun_out/ncu_fps.log 2>&1
tail -3 gpurun_out/ncu_fps.log
for v in v1 v2; do
  ncu -i gpurun_out/prof_r2_fps_$v.ncu-rep --page source --csv > gpurun_out/r2_fps_${v}_source.csv 2>/dev/null
  ncu -i gpurun_out/prof_r2_fps_$v.ncu-rep --page raw --csv > gpurun_out/r2_fps_${v}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r2_fps_*
