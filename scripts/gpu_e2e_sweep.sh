#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python - <<'PY'
import torch, time
x = torch.empty(512*1024*1024//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s: %.1f" % (5*x.numel()*4/1e9/(e0.elapsed_time(e1)/1e3)))
e0.record()
for _ in range(5): x.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H pinned GB/s: %.1f" % (5*x.numel()*4/1e9/(e0.elapsed_time(e1)/1e3)))
PY
for cfg in ${CFGS:-148:2}; do
  c=${cfg%%:*}; rest=${cfg#*:}; s=${rest%%:*}; g=${rest#*:}; [ "$g" = "$rest" ] && g=2,3,2,1
  timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk $c --e2e-streams $s --e2e-groups $g > gpurun_out/e2e_${c}_${s}.log 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_${c}_${s}.log").read().strip().splitlines()[-1])
print("chunk $c streams $s groups $g: e2e %.3e  (device value %.3e)"%(d["e2e"]["value"], d["value"]))
PY
done
