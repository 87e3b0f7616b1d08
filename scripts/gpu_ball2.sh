#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -p no:cacheprovider --timeout 300 -x tests/test_gpu_pointnet2.py -k "ball" > gpurun_out/t_ball.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/t_ball.log | cut -c1-300
NCU=$NCU bash scripts/gpu_ballpath.sh
