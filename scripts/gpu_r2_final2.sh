cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for f in test_gpu_pointops test_gpu_reference_live; do
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -W ignore > gpurun_out/sanitize_$f.log 2>&1
  echo "$f rc=$?"; grep -n "Invalid\|Device Frame\|passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_$f.log | head -8 | cut -c1-260
done
bash scripts/gpu_r2_final.sh
