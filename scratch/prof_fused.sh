#!/bin/bash
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass --launch-skip 3 -c 1 \
   -o gpurun_out/prof_r01_fused_v1 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_fused_v1.log 2>&1
tail -3 gpurun_out/prof_fused_v1.log
