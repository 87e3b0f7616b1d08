#!/bin/bash
# memcheck of the cluster-resident FPS kernel on the tiny / ragged cases
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_pointops.py -q -x -m gpu -p no:cacheprovider -W ignore -k "(fps or knn or ball or three or gather or interp) and (tiny or ragged or small or edge or empty or dups)" > gpurun_out/sanitize_fps.log 2>&1
echo rc=$?; grep -n "Invalid\|at 0x\|by thread\|in block\|Address\|passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_fps.log | head -40
