#!/bin/bash
# FPS bucket kernel: warps per cloud x batch size sweep (stage_ms.fps of bench.py)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CFGS=${CFGS:-1184:-14 1184:-12 2368:-14 2368:-12 2368:-11 4736:-12 4736:-11}
for cfg in $CFGS; do
  c=${cfg%%:*}; m=${cfg##*:}
  timeout -k 10 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --clouds $c --fps-mode=$m > gpurun_out/shape_${c}_${m}.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/shape_${c}_${m}.log").read().strip().splitlines()[-1])
    print("clouds $c mode $m: fps %.3f ms  step %.3f ms  value %.3e parity %s"%(d["stage_ms"]["fps"], d["ms_per_step"], d["value"], d.get("parity_ok")))
except Exception as e:
    print("clouds $c mode $m: FAILED", e); print(open("gpurun_out/shape_${c}_${m}.log").read()[-600:])
PY
done
