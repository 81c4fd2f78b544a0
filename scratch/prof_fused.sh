#!/bin/bash
cd /root/repo
TAG=${1:-v2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass --launch-skip 3 -c 1 \
   -o gpurun_out/prof_r01_fused_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_fused_$TAG.log 2>&1
tail -2 gpurun_out/prof_fused_$TAG.log | cut -c1-200
