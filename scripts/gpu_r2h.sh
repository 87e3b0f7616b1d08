#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -x -W ignore"
TAILN=25 run t_live $PT tests/test_gpu_reference_live.py
TAILN=10 run t_pointnet2 $PT tests/test_gpu_pointnet2.py
TMO=900 run model_parity python scripts/model_parity.py --out gpurun_out/model_parity.json --cases tseg
