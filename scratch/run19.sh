cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r19_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r19_pytest.log
timeout 900 python bench.py > gpurun_out/bench_r01_final3.json 2> gpurun_out/bench_r01_final3.err; echo "bench rc $?"; tail -2 gpurun_out/bench_r01_final3.err
python -c "
import json
j=json.loads(open('gpurun_out/bench_r01_final3.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['roofline']['frac'], j['gpu_launches'], j['roofline']['stages_ms']); print(j['e2e']['ms_per_step'], j['e2e']['breakdown_ms'], j['e2e']['float32_host_tables']['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_pass --launch-skip 3 -c 1 -o gpurun_out/prof_r01_fused_final3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_fused_final3.log 2>&1; tail -1 gpurun_out/prof_fused_final3.log | cut -c1-100
ABCDLS_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_final3.log 2>&1
tail -1 gpurun_out/ncu_final3.log | cut -c1-100
