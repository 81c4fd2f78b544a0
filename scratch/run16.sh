cd /root/repo
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py --big 1048576 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r16_bench8.json 2> gpurun_out/r16_bench8.err; echo "bench8 rc $?"
python -c "
import json
j=[json.loads(l) for l in open('gpurun_out/r16_bench8.json') if l.startswith('{')][-1]; print(j['n_gpus'], j['ms_per_step'], j['value'], j['roofline']['frac'], j['gpu_launches'], j['collectives'], j['roofline']['stages_ms']); print(j['e2e']['ms_per_step'], j['e2e']['breakdown_ms'], j['e2e']['float32_host_tables']['ms_per_step'])"
