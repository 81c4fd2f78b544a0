#!/bin/bash
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass --launch-skip 3 -c 1 \
   -o gpurun_out/prof_r01_fused_final -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_fused_final.log 2>&1
ABCDLS_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log | cut -c1-100
