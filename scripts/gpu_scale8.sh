#!/bin/bash
# 8-GPU weak scaling of the bench line, with and without NUMA binding (e2e is the number that moves)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
run() { n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 --no-extras --no-cpu-baseline "$@" 2>gpurun_out/scale_${n}_${tag}.err | grep '^{' | tail -1 > gpurun_out/scale_${n}_${tag}.json
  python -c "
import json
d = json.load(open('gpurun_out/scale_${n}_${tag}.json'))
print('N=$n $tag: value %.4e  e2e %.4e  ms %.2f  numa %s parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d.get('numa'), d['parity_ok']))
"
}
run 8 numa
run 8 nonuma --no-numa
run 4 numa
run 2 numa
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/scale_1_numa.json
python -c "
import json
d = json.load(open('gpurun_out/scale_1_numa.json')); print('N=1 numa: value %.4e e2e %.4e' % (d['value'], d['e2e']['value']))"
