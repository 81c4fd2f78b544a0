cd /root/repo
timeout 900 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r9_bench.err; python -c "
import json
j=json.loads(open('gpurun_out/r9_bench.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['roofline']['frac'], j['gpu_launches']); print(json.dumps(j['e2e'], indent=1)); print(j['cpu_baseline'])"
