#!/bin/bash
# Quick GPU pass: FPS parity + bench (+ optional ncu of one kernel regex via $NCU_K).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-1000; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run t_fps $PT tests/test_gpu_pointops.py -k "fps"
run bench1184 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clouds 1184
run bench592 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clouds 592
if [ -n "$NCU_K" ]; then
  TMO=900 run ncu_full ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-10} -c ${NCU_C:-2} -o gpurun_out/prof_quick -f python bench.py --clouds 296 --steps 1 --warmup 3 --no-cpu-baseline
fi
