cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r10_pytest.log
for lay in row col; do
timeout 900 python bench.py --param-layout $lay --no-cpu --no-e2e-f32 > gpurun_out/r10_bench_$lay.json 2> gpurun_out/r10_bench_$lay.err; echo "bench $lay rc $?"; tail -3 gpurun_out/r10_bench_$lay.err; python -c "
import json
j=json.loads(open('gpurun_out/r10_bench_$lay.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['roofline']['frac'], j['gpu_launches'], j['roofline']['stages_ms']); print(j['e2e']['ms_per_step'], j['e2e']['breakdown_ms'])"
done
