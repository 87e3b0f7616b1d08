#!/bin/bash
# Full GPU pass: every -m gpu test (grouped so a hung kernel only loses its group), smoke, bench, profiles.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-r1}
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-500; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run t_pointops  $PT tests/test_gpu_pointops.py
run t_pointnet2 $PT tests/test_gpu_pointnet2.py
run t_callers $PT tests/test_gpu_callers.py
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 10 --warmup 3
run bench592 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clouds 592
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
TMO=900 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
TMO=1200 run ncu_full ncu --set full --clock-control none --import-source on -k regex:"fps_bucket_kernel|sa_mlp_tc|ball_query_kernel|fps_bucket_sort" -s 12 -c 4 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$TAG.txt 2>&1
