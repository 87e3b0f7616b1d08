#!/bin/bash
# round 2, call B: new pw engine + layout-driven norm rounding: live-reference tests, the whole GPU suite, model parity, op bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -x"
TAILN=40 run t_live $PT tests/test_gpu_reference_live.py
TAILN=15 run t_pointnet2 $PT tests/test_gpu_pointnet2.py
run t_pointops  $PT tests/test_gpu_pointops.py
run t_callers $PT tests/test_gpu_callers.py
run smoke python __graft_entry__.py smoke
TMO=900 run model_parity python scripts/model_parity.py --out gpurun_out/model_parity.json
