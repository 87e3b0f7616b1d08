#!/bin/bash
# every -m gpu test exactly as the driver runs them, then smoke()
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 600 -W ignore > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/t_all.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
