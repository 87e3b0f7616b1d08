#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run t_pn2 $PT tests/test_gpu_pointnet2.py
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run bench296 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clouds 296
nproc; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count()); print(open('/sys/fs/cgroup/cpu.max').read())"
run bench_ref python bench.py --impl reference --steps 1 --warmup 0
