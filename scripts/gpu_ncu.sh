#!/bin/bash
# One ncu --set full capture of kernels matching $NCU_K in the bench step -> gpurun_out/prof_$TAG.ncu-rep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-3} -c ${NCU_C:-1} \
  -o gpurun_out/prof_${TAG:-x} -f python bench.py --clouds ${CLOUDS:-592} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG:-x}.log 2>&1
echo rc=$?; tail -n 2 gpurun_out/ncu_${TAG:-x}.log | cut -c1-300
