#!/bin/bash
# Third GPU pass: bucket-pruned FPS parity + timing, ncu of the bucket kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-1200; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run t_fps   $PT tests/test_gpu_pointops.py -k "fps"
run t_sa    $PT tests/test_gpu_pointnet2.py -k "set_abstraction or fused or farthest"
for m in 0 2; do
  run bench_fps$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fps-mode $m
done
run bench_c592 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clouds 592
TMO=900 run ncu_full ncu --set full --clock-control none --import-source on -k regex:"fps_bucket" -s 10 -c 2 -o gpurun_out/prof_r1b -f python bench.py --clouds 296 --steps 1 --warmup 3 --no-cpu-baseline
