#!/bin/bash
# Launch-duration list of the bench step (ncu, cold-cache, serialised) -> gpurun_out/launches_$TAG.csv
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-x}
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_S:-60} -c ${NCU_C:-60} --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/launches_$TAG.log 2>&1
echo rc=$?
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e6:8.3f} ms")
PY
