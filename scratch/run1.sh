#!/bin/bash
# first GPU run of the one-pass path: tests, then a short bench
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -15
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_fused1.json 2> gpurun_out/bench_fused1.err
tail -3 gpurun_out/bench_fused1.err
cat gpurun_out/bench_fused1.json
