#!/bin/bash
# fused PointTransformerLayer / TransitionDown + subtract-first BatchNorm: tests, model parity beside the reference
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_reference_live.py tests/test_gpu_pointnet2.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_blocks_tests.log
cat gpurun_out/r2_blocks_tests.log
timeout 900 python scripts/model_parity.py --cases pp,tseg,tgni --out gpurun_out/r2_model_parity_bn2.json > gpurun_out/r2_model_parity_bn2.log 2>&1
grep -n "worst_max_rel\|ours_vs_fp64\|reference_vs_fp64\|\"pass\"\|reference\"\|b200" gpurun_out/r2_model_parity_bn2.log
