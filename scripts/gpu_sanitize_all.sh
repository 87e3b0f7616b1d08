#!/bin/bash
# memcheck (caching allocator off, so every tensor is its own allocation) over the small-case test files
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
for f in test_gpu_blocks test_gpu_clustering test_gpu_crops test_gpu_callers test_gpu_pointnet2; do
  timeout 420 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -W ignore > gpurun_out/sanitize_$f.log 2>&1
  echo "$f rc=$?"; grep -n "Invalid\|Device Frame\|passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_$f.log | head -8 | cut -c1-260
done
