#!/bin/bash
# end-to-end pipeline sweep (round 2): chunk x streams, NUMA-bound; prints value / e2e per config
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "148 2" "148 3" "148 4" "296 2" "296 3" "74 4" "74 6" "148 2 nonuma"; do
  set -- $cfg
  extra=""; [ "$3" = "nonuma" ] && extra="--no-numa"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --e2e-chunk $1 --e2e-streams $2 $extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('chunk $1 streams $2 $3: value %.3e e2e %.3e ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value']))
"
done 2>&1 | tee gpurun_out/e2e_sweep2.log
python scripts/e2e_timeline.py > gpurun_out/e2e_timeline2.log 2>&1; tail -30 gpurun_out/e2e_timeline2.log
