#!/bin/bash
# Verification pass: every -m gpu test, smoke, default bench (with CPU baseline), reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-300}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-400; }
run t_gpu python -m pytest -q -p no:cacheprovider --timeout 300 -m gpu tests
run smoke python __graft_entry__.py smoke
run bench python bench.py
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value=%.3e ms/step=%.3f'%(d['value'],d['ms_per_step']), d['stage_ms'], 'e2e=%.3e'%d['e2e']['value'], 'cpu=%.3e'%d['cpu_baseline']['value'], d['clocks'], d['parity_ok'], 'launches', d['gpu_launches'])
print({k: d['roofline'][k] for k in ('achieved','frac','traffic','dram_achieved_gbs','dram_frac')})
PY
