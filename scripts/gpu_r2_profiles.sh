#!/bin/bash
# round 2 profiles: launch lists (bench step, pointnet_pp forward in train mode) + one full capture of the layer kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_1184clouds.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_list_bench.log 2>&1
echo "bench list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "measured_forward/" --csv --log-file gpurun_out/r2_launches_pointnet_pp_trainbn.csv python scripts/ncu_pp_driver.py > gpurun_out/ncu_list_pp.log 2>&1
echo "pp list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pw_layer_kernel" -s 20 -c 3 -o gpurun_out/prof_r2_pw -f python scripts/ncu_pp_driver.py > gpurun_out/ncu_pw.log 2>&1
echo "pw full rc=$?"
ncu -i gpurun_out/prof_r2_pw.ncu-rep --page raw --csv > gpurun_out/r2_ncu_pw_layer_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_grid_query_kernel" -c 1 -o gpurun_out/prof_r2_knn -f python -c "
import sys, torch
sys.path.insert(0, '.')
from toothgroupnetwork_b200 import clouds, pointops
x = clouds.dental_arch(24000, 0)[0].cuda(); o = torch.tensor([24000], dtype=torch.int32).cuda()
pointops.knn_packed(36, x, x, o, o); torch.cuda.synchronize()
" > gpurun_out/ncu_knn.log 2>&1
ncu -i gpurun_out/prof_r2_knn.ncu-rep --page raw --csv > gpurun_out/r2_ncu_knn_grid_raw.csv 2>/dev/null
ls -la gpurun_out/r2_*
