cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r12_pytest.log
timeout 900 python bench.py > gpurun_out/bench_r01_final2.json 2> gpurun_out/bench_r01_final2.err; echo "bench rc $?"; tail -2 gpurun_out/bench_r01_final2.err
python -c "
import json
j=json.loads(open('gpurun_out/bench_r01_final2.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['roofline']['frac'], j['gpu_launches'], j['roofline']['stages_ms']); print(j['e2e']['ms_per_step'], j['e2e']['breakdown_ms'], j['e2e']['float32_host_tables']['ms_per_step'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_final2_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r01_final2_ref.json
ABCDLS_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_final2.log 2>&1
tail -1 gpurun_out/ncu_final2.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small_columns --launch-skip 2 -c 1 -o gpurun_out/prof_r01_small -f python bench.py --workload cv4abc > gpurun_out/prof_small.log 2>&1; tail -1 gpurun_out/prof_small.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gather_param_rm|k_normal_partial" --launch-skip 6 -c 3 -o gpurun_out/prof_r01_tail -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_tail.log 2>&1; tail -1 gpurun_out/prof_tail.log | cut -c1-100
