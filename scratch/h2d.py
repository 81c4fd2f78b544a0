import torch, time
dev = torch.device("cuda", 0)
n = 1 << 30  # 8 GB of fp64 = 1G doubles? use 2^29 doubles = 4 GB
h = torch.empty(1 << 29, dtype=torch.float64, pin_memory=True)
h.fill_(1.0)
d = torch.empty_like(h, device=dev)
for nstreams in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    chunks_h = h.chunk(nstreams); chunks_d = d.chunk(nstreams)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for s, a, b in zip(streams, chunks_h, chunks_d):
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(nstreams, "streams:", h.numel() * 8 / dt / 1e9, "GB/s")
# D2H
torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("d2h", h.numel() * 8 / (time.perf_counter() - t0) / 1e9, "GB/s")
