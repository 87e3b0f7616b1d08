#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -x -W ignore"
TAILN=25 run t_crops $PT tests/test_gpu_crops.py
TAILN=8 run t_pointnet2 $PT tests/test_gpu_pointnet2.py
TMO=900 run op_bench python scripts/op_bench.py --out gpurun_out/op_bench.json --sections knn
TMO=900 TAILN=40 run e2e_sweep bash scripts/gpu_e2e_sweep2.sh
