cd /root/repo
timeout 900 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -2 gpurun_out/bench_r01_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_final_ref.json 2>/dev/null
bash scratch/prof_final.sh
