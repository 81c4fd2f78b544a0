cd /root/repo
N=$1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py --big 1048576 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -7
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['n_gpus'], 'ms/step', j['ms_per_step'], 'host', j['host_enqueue_ms'], 'graph', j['cuda_graph'], j['roofline']['stages_ms'], 'miss', j['fast_path_misses'])
    else: print(l.rstrip()[:300])
"
