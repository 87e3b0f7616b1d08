#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -p no:cacheprovider --timeout 600 -x -W ignore"
TAILN=12 run diag_fp python scripts/diag_fp_bwd.py
TAILN=25 run t_live $PT tests/test_gpu_reference_live.py -k "feature_propagation"
TMO=900 TAILN=40 run e2e_sweep3 bash scripts/gpu_e2e_sweep3.sh
