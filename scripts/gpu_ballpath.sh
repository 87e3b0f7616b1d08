#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for bp in 4 8; do
  timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ball-path $bp > gpurun_out/ballpath_$bp.log 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/ballpath_$bp.log").read().strip().splitlines()[-1])
print("ball-path $bp:", d["stage_ms"], "step %.3f"%d["ms_per_step"], d.get("parity_ok"))
PY
done
if [ -n "$NCU" ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"ball_query_grid" -s 3 -c 1 -o gpurun_out/prof_ballgrid -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ball-path 8 --clouds 592 > gpurun_out/ncu_ballgrid.log 2>&1; echo ncu rc=$?
fi
