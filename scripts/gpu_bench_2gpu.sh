#!/bin/bash
# the driver's N = 2 launch of both arms, default flags (extras and cpu_baseline on rank 0)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.out 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; grep '^{' gpurun_out/bench_2gpu.out | tail -1 > gpurun_out/bench_2gpu.json; wc -c gpurun_out/bench_2gpu.json; tail -c 400 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_2gpu.json'))
print(d['numa']); print({k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step', 'parity_ok', 'gpu_launches', 'extras_error')}, d['e2e']['value'], list(d.get('next_rows', {}).keys()), d.get('cpu_baseline', {}).get('value'))
print({k: (v.get('error') or v.get('unavailable') or 'ok') for k, v in d.get('next_rows', {}).items()})
PY
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1; echo "CVD=$CUDA_VISIBLE_DEVICES"

