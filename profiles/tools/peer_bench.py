"""Latency of the NVLink mailbox exchanges in isolation (ranks in lockstep), next to NCCL, launched with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        profiles/tools/peer_bench.py

Prints us per exchange (CUDA events, max over ranks) for the message shapes one abc() call uses."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from abc_dls_b200.comm import Comm  # noqa: E402


def timeit(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e3 / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm(None, dev)
    assert comm.peer is not None, "mailboxes unavailable"
    out = {}
    for name, n, dt in (("allreduce f64 x128", 128, torch.float64), ("allreduce f64 x1024 (sample est)", 1024, torch.float64),
                        ("allreduce i64 x16480 (pass counts + median histogram)", 16480, torch.int64),
                        ("allreduce i64 x24608 (fine sample counts)", 24608, torch.int64)):
        t = torch.ones(n, dtype=dt, device=dev)
        out["peer " + name] = timeit(lambda: comm.allreduce_sum(t))
        out["nccl " + name] = timeit(lambda: dist.all_reduce(t))
        t.fill_(1)
    src = torch.ones(8193, dtype=torch.float64, device=dev)
    out["peer allgather 8193 f64 (ds block)"] = timeit(lambda: comm.allgather(src))
    one = torch.ones(1, dtype=torch.int64, device=dev)
    out["peer allgather 1 i64 (cap rule)"] = timeit(lambda: comm.allgather(one))
    vals = torch.ones(16 * 8192, dtype=torch.float64, device=dev)
    cnt = torch.full((16,), 300, dtype=torch.int32, device=dev)
    out["peer allgather_ragged 16 x 300 (refinement block)"] = timeit(lambda: comm.allgather_ragged(vals, cnt))
    # a dependent chain of small kernels between exchanges, as in a call (launch gaps included): eager vs one CUDA graph
    x = torch.ones(1024, dtype=torch.float64, device=dev)

    def chain():
        for _ in range(4):
            x.mul_(1.0)
            comm.allreduce_sum(x)
    out["eager: 4 x (tiny kernel + allreduce x1024)"] = timeit(chain, 50)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.graph(g, stream=s):
            chain()
    torch.cuda.synchronize()
    dist.barrier()
    out["graph: 4 x (tiny kernel + allreduce x1024)"] = timeit(g.replay, 50)
    if rank == 0:
        print(f"world {world}")
        for k, v in out.items():
            print(f"  {v:8.2f} us  {k}")
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
