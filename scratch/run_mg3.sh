cd /root/repo
timeout 200 python -m pytest tests -m gpu -x -q -k "graph or fused_one_pass_parity or postpr" 2>&1 | tail -3
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['n_gpus'], j['ms_per_step'], j['host_enqueue_ms'], j['cuda_graph'], j['roofline']['stages_ms'], j['fast_path_misses'])
    else: print(l.rstrip()[:300])
"
