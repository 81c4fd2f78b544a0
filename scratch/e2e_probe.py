import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abc_dls_b200 import synth
from abc_dls_b200.engine import AbcEngine
dev = torch.device("cuda", 0)
eng = AbcEngine(dev)
N, d, P = 100000000, 16, 16
par, ss, tgt = synth.abc_tables(N, d, P, dev)
h_par = torch.empty((P, N), dtype=torch.float64, pin_memory=True); h_par.copy_(par)
h_ss = torch.empty((d, N), dtype=torch.float64, pin_memory=True); h_ss.copy_(ss)
del par
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); dss = h_ss.to(dev, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    eng.timers = {}
    r = eng.abc(tgt, h_par, dss, 0.01, "loclinear"); torch.cuda.synchronize(); t2 = time.perf_counter()
    tm = {k: round(sum(v) / len(v), 3) for k, v in eng.timer_ms().items()}
    eng.timers = None
    out = r.adj_values.contiguous().cpu(); t3 = time.perf_counter()
    print(f"h2d {1e3*(t1-t0):.1f} ms  abc {1e3*(t2-t1):.1f} ms  d2h {1e3*(t3-t2):.1f} ms", tm)
    del dss
