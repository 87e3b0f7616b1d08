#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() {
  extra=""; [ "$3" != "0" ] && extra="--e2e-fps-mode $3"; [ "$4" != "0" ] && extra="$extra --e2e-sa-engine $4"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-numa --e2e-chunk $1 --e2e-streams $2 $extra $5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('chunk $1 streams $2 fps $3 sa $4 $5: value %.3e e2e %.3e ratio %.3f parity %s' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value'], d['parity_ok']))
"
}
{
one 148 2 0 0
one 148 4 -14 5
one 148 4 0 5
one 148 8 -14 5
one 74 8 -14 5
one 148 4 -18 0
} 2>&1 | tee gpurun_out/e2e_sweep5.log
python scripts/e2e_timeline.py 1184 148 4 > gpurun_out/e2e_timeline5.log 2>&1; tail -12 gpurun_out/e2e_timeline5.log
