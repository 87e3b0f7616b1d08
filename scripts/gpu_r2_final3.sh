#!/bin/bash
# ncu full capture of the transformer-layer passes and the DBSCAN neighbour passes, then the whole -m gpu suite and smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measured/" -k regex:"pt_layer_kernel|db_neighbour_kernel" -o gpurun_out/r2_ncu_blocks_dbscan python scripts/ncu_blocks_driver.py > gpurun_out/ncu_blocks.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_blocks.log; ls -la gpurun_out/r2_ncu_blocks_dbscan.ncu-rep
timeout 1800 python -m pytest tests/ -q -m gpu -p no:cacheprovider --timeout 600 -W ignore > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/t_all.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
