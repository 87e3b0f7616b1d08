#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -p no:cacheprovider --timeout 300 -x -W ignore"
TAILN=15 run t_sa $PT tests/test_gpu_pointnet2.py -k "fused_sa_matches_oracle or interpolation_forward_backward"
for e in 2 4; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --sa-engine $e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('engine $e: value %.3e stage_ms %s parity %s e2e %.3e' % (d['value'], d['stage_ms'], d['parity_ok'], d['e2e']['value']))
"
done 2>&1 | tee gpurun_out/engine_cmp.log
