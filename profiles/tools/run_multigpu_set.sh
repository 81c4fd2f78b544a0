# the N-GPU measurement set behind profiles/r02_*_${NG}gpu_* and BASELINE.md (run under gpurun --gpus $NG from the repo root):
#   NG=8 bash profiles/tools/run_multigpu_set.sh
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
run() { n=$1; shift; timeout 400 $T bench.py --gpus $NG "$@" 2> gpurun_out/b${NG}.err | grep "^{" > gpurun_out/r02_bench_${n}_${NG}gpu.json; python - <<PY
import json
j=json.loads(open('gpurun_out/r02_bench_${n}_${NG}gpu.json').read().strip().splitlines()[-1])
print(j['n_gpus'], '${n}', round(j['ms_per_step'],4), '%.4g'%j['value'], j['roofline']['stages_ms'], j.get('collectives'), (j.get('e2e') or {}).get('ms_per_step'), j.get('host_profile_us'))
if 'sweep' in j: print([(e['rows'], round(e['ms_per_step'],4)) for e in j['sweep']])
PY
}
run cfg5 --steps 20 --warmup 5 --no-cpu
if [ "${FULL:-1}" = "1" ]; then
python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multigpu_${NG}gpu.log
run cfg3 --config cfg3 --steps 20 --warmup 5 --no-cpu --no-e2e
run cfg2_postpr --workload postpr --steps 20 --warmup 5 --no-cpu
run cfg4_smc --workload smc --steps 3 --warmup 1 --no-cpu
run sweep --workload sweep --steps 10 --warmup 3 --no-cpu
fi
tail -2 gpurun_out/b${NG}.err
