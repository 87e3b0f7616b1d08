#!/bin/bash
# round 2, call A: model-level parity (reference models on both operator sets) + operator timings beside the reference GPU path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 5 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run model_parity python scripts/model_parity.py --out gpurun_out/model_parity.json
TMO=900 run op_bench python scripts/op_bench.py --out gpurun_out/op_bench.json
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt 2>&1
numactl -H >> gpurun_out/lscpu.txt 2>&1
ls /sys/bus/pci/devices/*/numa_node | head -3 >> gpurun_out/lscpu.txt 2>&1
