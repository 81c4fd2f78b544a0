cd /root/repo
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r14_bench2.json 2> gpurun_out/r14_bench2.err; echo "bench2 rc $?"; tail -3 gpurun_out/r14_bench2.err | cut -c1-300
python -c "
import json
j=[json.loads(l) for l in open('gpurun_out/r14_bench2.json') if l.startswith('{')][-1]; print(j['n_gpus'], j['ms_per_step'], j['roofline']['frac'], j['gpu_launches'], j['collectives']); print(j['e2e']['ms_per_step'], j['e2e']['breakdown_ms'], j['e2e']['float32_host_tables']['ms_per_step']); print(j['cpu_baseline']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-160
