cd /root/repo
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json; echo
timeout 900 python bench.py > gpurun_out/bench_r01_b.json 2> gpurun_out/bench_r01_b.err; tail -3 gpurun_out/bench_r01_b.err; cat gpurun_out/bench_r01_b.json
