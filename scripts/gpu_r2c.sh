#!/bin/bash
# round 2, call C: every GPU test, smoke, the default bench line + reference arm, operator timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -x -W ignore"
TAILN=25 run t_live $PT tests/test_gpu_reference_live.py
TAILN=15 run t_pointnet2 $PT tests/test_gpu_pointnet2.py
TAILN=15 run t_pointops  $PT tests/test_gpu_pointops.py
run t_callers $PT tests/test_gpu_callers.py
run t_crops $PT tests/test_gpu_crops.py
run smoke python __graft_entry__.py smoke
TAILN=3 run bench python bench.py --steps 10 --warmup 3
TAILN=3 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
TMO=900 run op_bench python scripts/op_bench.py --out gpurun_out/op_bench.json --sections fps,knn,ball,sa,fp
TMO=900 run model_parity python scripts/model_parity.py --out gpurun_out/model_parity.json
