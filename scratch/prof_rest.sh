#!/bin/bash
cd /root/repo
TAG=${1:-v1}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_normal_partial|k_gather_cand|k_adjust|k_residual|k_cand_exact|k_sample_cols|k_sample_qdist|k_refine_hist|k_refine_gather|k_recentre_log|k_col_stats|k_solve' --launch-skip 60 -c 20 \
   -o gpurun_out/prof_r01_rest_$TAG -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_rest_$TAG.log 2>&1
tail -2 gpurun_out/prof_rest_$TAG.log | cut -c1-200
