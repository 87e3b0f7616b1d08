bash scripts/gpu_tgn_list.sh nograd
timeout 800 python scripts/model_parity.py --cases tgni --out gpurun_out/r2_model_parity_tgni.json 2>&1 | grep -n "worst_max_rel\|ours_vs_fp64\|reference_vs_fp64\|\"pass\"\|reference\"\|b200"
