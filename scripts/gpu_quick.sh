#!/bin/bash
# dev helper: parity checks + kNN timing on the GPU box
python scripts/gpu_check_dm.py > gpurun_out/check_dm.log 2>&1; echo "dm exit=$?"
grep -E "knn|K indptr|subspace|run_diffusion|TFLOP" gpurun_out/check_dm.log
python scripts/prof_knn.py --reps 3
python scripts/prof_knn.py --n 1000000 --d 100 --reps 2
python scripts/prof_knn.py --n 100000 --d 50 --reps 3
