#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300 -x"
run t_ball $PT tests/test_gpu_pointnet2.py -k "ball"
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value=%.3e ms/step=%.3f'%(d['value'],d['ms_per_step']), d['stage_ms'], 'e2e=%.3e'%d['e2e']['value'], d.get('parity_ok'))
PY
