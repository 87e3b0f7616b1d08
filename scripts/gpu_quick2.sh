#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-420}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300"
run t_sa $PT tests/test_gpu_pointnet2.py -k "set_abstraction or fused or group_all or unsupported or training"
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value=%.3e ms/step=%.3f'%(d['value'],d['ms_per_step']), d['stage_ms'], 'e2e=%.3e'%d['e2e']['value'])
PY
if [ -n "$NCU_K" ]; then
  TMO=900 run ncu_full ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-5} -c ${NCU_C:-1} -o gpurun_out/prof_quick -f python bench.py --clouds 296 --steps 1 --warmup 3 --no-cpu-baseline
fi
