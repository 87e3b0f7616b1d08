#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "148 4" "148 2" "296 2" "74 4" "592 1"; do
  set -- $cfg
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk $1 --e2e-streams $2 > gpurun_out/e2e_$1_$2.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_$1_$2.log').read().strip().splitlines()[-1])
print('chunk=$1 streams=$2 value=%.3e e2e=%.3e ms/step=%.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']))
PY
done
