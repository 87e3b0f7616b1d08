#!/bin/bash
# Second GPU pass: re-check FPS / FP parity, compare FPS shapes, launch list + ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-1500; }
PT="python -m pytest -q -p no:cacheprovider --timeout 180"
run t_fps   $PT tests/test_gpu_pointops.py -k "fps"
run t_fp    $PT tests/test_gpu_pointnet2.py -k "feature_prop"
for m in 0 2 4 8 202 204 208; do
  run bench_fps$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fps-mode $m
done
run bench_fp32mlp python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sa-engine 1
TMO=600 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --clouds 296 --steps 2 --warmup 3 --no-cpu-baseline
TMO=900 run ncu_full ncu --set full --clock-control none --import-source on -k regex:"fps_resident|sa_mlp_tc|ball_query" -s 15 -c 3 -o gpurun_out/prof_r1 -f python bench.py --clouds 148 --steps 1 --warmup 3 --no-cpu-baseline
ls -la gpurun_out | head -40
