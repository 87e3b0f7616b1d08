#!/bin/bash
# fixtures-based block / clustering tests, both pt_layer schedules, tgnet no-grad step with GPU clustering, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_clustering.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_blocks2_tests.log
timeout 600 python scripts/pt_layer_bench.py 2>&1 | grep -v Warn | tee gpurun_out/r2_pt_layer_bench.log
timeout 800 python scripts/model_parity.py --cases tgni --out gpurun_out/r2_model_parity_tgni.json 2>&1 | grep -n "worst_max_rel\|ours_vs_fp64\|reference_vs_fp64\|\"pass\"\|reference\"\|b200"
