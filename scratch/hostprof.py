import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abc_dls_b200 import synth
from abc_dls_b200.engine import AbcEngine
dev = torch.device("cuda", 0)
eng = AbcEngine(dev)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
par, ss, tgt = synth.abc_tables(N, 16, 16, dev)
for _ in range(3):
    eng.abc(tgt, par, ss, 0.01, "loclinear")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    eng.abc(tgt, par, ss, 0.01, "loclinear")
torch.cuda.synchronize()
print("ms/call", (time.perf_counter() - t0) / 5 * 1e3, "host enqueue", eng.host_enqueue_ms[-5:])
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    eng.abc(tgt, par, ss, 0.01, "loclinear")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
