#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:"sa_mlp_tcw" -s 4 -c 1 -o gpurun_out/prof_tcw -f python scripts/c2_real_sa1.py 148 > gpurun_out/ncu_tcw.log 2>&1
echo rc=$?
