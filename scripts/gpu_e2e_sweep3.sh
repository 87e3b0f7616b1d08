#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "148 4 -14" "148 8 -14" "74 8 -14" "296 4 -14" "148 4 -18" "148 8 -18" "74 16 -14" "148 3 0"; do
  set -- $cfg
  extra=""; [ "$3" != "0" ] && extra="--e2e-fps-mode $3"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-numa --e2e-chunk $1 --e2e-streams $2 $extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('chunk $1 streams $2 fps $3: value %.3e e2e %.3e ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value']))
"
done 2>&1 | tee gpurun_out/e2e_sweep3.log
