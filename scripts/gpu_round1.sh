#!/bin/bash
# First GPU pass: parity tests in isolated groups (a hung kernel only loses its group), smoke, bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 6 gpurun_out/$name.log; }
PT="python -m pytest -q -p no:cacheprovider --timeout 180"
run t_fps   $PT tests/test_gpu_pointops.py -k "fps"
run t_knn   $PT tests/test_gpu_pointops.py -k "knn"
run t_gath  $PT tests/test_gpu_pointops.py -k "not fps and not knn"
run t_ball  $PT tests/test_gpu_pointnet2.py -k "ball or three_nn or feature_prop or farthest"
run t_sa    $PT tests/test_gpu_pointnet2.py -k "not ball and not three_nn and not feature_prop and not farthest"
run smoke   python __graft_entry__.py smoke
TMO=600 run bench python bench.py --steps 5 --warmup 3
