#!/bin/bash
# end-of-round pass: every -m gpu test as the driver runs them, smoke, the default bench line, the reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 600 -W ignore > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/t_all.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.err; head -c 1500 gpurun_out/bench_default.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; head -c 800 gpurun_out/bench_reference.json; echo
