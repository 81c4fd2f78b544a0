cd /root/repo
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2 | cut -c1-1800
echo "---- graph collectives"
ABCDLS_GRAPH_COLLECTIVES=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4 | cut -c1-1800
