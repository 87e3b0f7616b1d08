#!/bin/bash
# GPU DBSCAN / get_clustering_labels: parity with scikit-learn and timing beside it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clustering.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_cluster_tests.log
timeout 600 python scripts/cluster_bench.py 2>&1 | grep -v Warning | tee gpurun_out/r2_cluster_bench.log
