cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r8_pytest.log
timeout 600 python scratch/e2e_probe2.py > gpurun_out/r8_probe.log 2>&1; echo "probe rc $?"; cat gpurun_out/r8_probe.log | tail -25
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; echo "bench rc $?"; python -c "
import json
j=json.loads(open('gpurun_out/r8_bench.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['roofline']['frac'], j['gpu_launches'], j['roofline']['stages_ms'])"
