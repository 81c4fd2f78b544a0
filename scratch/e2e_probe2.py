"""Where does the end-to-end time go?  Events around H2D / abc (pinned param, graph replay) / D2H at the bench shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abc_dls_b200 import synth
from abc_dls_b200.engine import AbcEngine
dev = torch.device("cuda", 0)
eng = AbcEngine(dev)
N, d, P = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000, 16, 16
par, ss, tgt = synth.abc_tables(N, d, P, dev)
h_par = torch.empty((P, N), dtype=torch.float64, pin_memory=True); h_par.copy_(par)
h_ss = torch.empty((d, N), dtype=torch.float64, pin_memory=True); h_ss.copy_(ss)
h_ss32 = torch.empty((d, N), dtype=torch.float32, pin_memory=True); h_ss32.copy_(ss)
h_out = torch.empty((P, N // 100 + 1), dtype=torch.float64, pin_memory=True)
del ss
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
for rep in range(5):
    e = [ev() for _ in range(5)]
    t0 = time.perf_counter()
    e[0].record()
    dss = h_ss.to(dev, non_blocking=True)
    e[1].record()
    r = eng.abc(tgt, h_par, dss, 0.01, "loclinear")
    e[2].record()
    vals = r.adj_values.t()
    h_out[:, :vals.shape[1]].copy_(vals, non_blocking=True)
    e[3].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"rep {rep}: wall {wall:.1f}  h2d {e[0].elapsed_time(e[1]):.1f}  abc(pinned par) {e[1].elapsed_time(e[2]):.1f}  d2h {e[2].elapsed_time(e[3]):.1f}  ptr {dss.data_ptr():x} graphs {len(eng._graphs)}", flush=True)
    del dss, r, vals
# param on the device instead of pinned
for rep in range(3):
    e = [ev() for _ in range(3)]
    dss = h_ss.to(dev, non_blocking=True)
    e[0].record()
    r = eng.abc(tgt, par, dss, 0.01, "loclinear")
    e[1].record()
    torch.cuda.synchronize()
    print(f"device par: abc {e[0].elapsed_time(e[1]):.2f}", flush=True)
    del dss, r
# float32 host table, widened on the device
for rep in range(4):
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record()
    dss = h_ss32.to(dev, non_blocking=True).to(torch.float64)
    e[1].record()
    r = eng.abc(tgt, h_par, dss, 0.01, "loclinear")
    e[2].record()
    vals = r.adj_values.t()
    h_out[:, :vals.shape[1]].copy_(vals, non_blocking=True)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"f32 host rep {rep}: wall {wall:.1f}  h2d+widen {e[0].elapsed_time(e[1]):.1f}  abc {e[1].elapsed_time(e[2]):.1f}  ptr {dss.data_ptr():x}", flush=True)
    del dss, r, vals
# H2D in chunks on several streams
for nchunk in (1, 2, 4):
    dst = torch.empty((d, N), dtype=torch.float64, device=dev)
    streams = [torch.cuda.Stream(dev) for _ in range(nchunk)]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for s_, a, b in zip(streams, h_ss.chunk(nchunk, 0), dst.chunk(nchunk, 0)):
            with torch.cuda.stream(s_):
                b.copy_(a, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"h2d {nchunk} streams: {h_ss.numel() * 8 / dt / 1e9:.1f} GB/s", flush=True)
    del dst
