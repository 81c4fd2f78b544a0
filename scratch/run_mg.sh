cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py --big 2097152 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | cut -c1-1500
