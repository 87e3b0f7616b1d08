#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-700; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300 -x -W ignore"
TAILN=12 run t_fps $PT tests/test_gpu_pointops.py -k "fps_matches_oracle"
TAILN=12 run fps_modes python scripts/fps_modes_bench.py
