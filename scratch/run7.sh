cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['host_enqueue_ms'], j['roofline']['stages_ms'], j['fast_path_misses'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_normal_partial|k_residual_stats|k_gather_cand|k_refine_hist|k_sample_cols' --launch-skip 15 -c 6 -o gpurun_out/prof_r01_rest_v2 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_rest_v2.log 2>&1
tail -1 gpurun_out/prof_rest_v2.log | cut -c1-80
