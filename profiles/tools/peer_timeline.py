"""Reads gpurun_out/peer_trace_r*.npy (bench.py under ABCDLS_PEER_TRACE=1) and prints, per exchange of a call
(averaged over the last calls): what the ranks spent before publishing (local copy + grid barrier), waiting for the
slowest peer's flag, reading the peers' data, and computing until the next exchange.  us.

    python profiles/tools/peer_timeline.py [exchanges_per_call] [calls]
"""
import glob
import sys

import numpy as np

per = int(sys.argv[1]) if len(sys.argv) > 1 else 12
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tr = [np.load(f).astype(np.int64) for f in sorted(glob.glob("gpurun_out/peer_trace_r*.npy"))]
n = min(len(t) for t in tr)
tr = [t[-(per * calls):] for t in tr]
W = len(tr)
A = np.stack(tr)                       # (world, per * calls, 4)
A = A.reshape(W, calls, per, 4)
pre = (A[..., 1] - A[..., 0]) / 1e3    # entry -> published
wait = (A[..., 2] - A[..., 1]) / 1e3   # published -> all flags seen
read = (A[..., 3] - A[..., 2]) / 1e3   # flags -> exit
nxt = np.empty_like(pre)
nxt[:, :, :-1] = (A[:, :, 1:, 0] - A[:, :, :-1, 3]) / 1e3      # exit -> next entry (compute + launch gaps)
nxt[:, :-1, -1] = (A[:, 1:, 0, 0] - A[:, :-1, -1, 3]) / 1e3    # ... across the call boundary (incl. host)
nxt[:, -1, -1] = np.nan
print(f"world {W}, {calls} calls x {per} exchanges; us, mean over calls; min / max over ranks")
print(" #   pre(min max)   wait(min  max)   read(min max)   compute-to-next(min max)")
tot = np.zeros(4)
for e in range(per):
    f = lambda x: (np.nanmean(x[:, :, e], axis=1).min(), np.nanmean(x[:, :, e], axis=1).max())
    a, b, c, d = f(pre), f(wait), f(read), f(nxt)
    tot += [a[1], b[0], c[1], d[1]]
    print(f"{e:2d}  {a[0]:5.1f} {a[1]:5.1f}   {b[0]:6.1f} {b[1]:6.1f}   {c[0]:5.1f} {c[1]:5.1f}   {d[0]:7.1f} {d[1]:7.1f}")
call = (A[:, 1:, 0, 0] - A[:, :-1, 0, 0]).mean() / 1e3
print(f"call period {call:.1f} us; sum of pre(max) {tot[0]:.1f}, wait(min = protocol floor) {tot[1]:.1f}, read(max) {tot[2]:.1f}")
print(f"wait beyond the floor (skew), mean over ranks: {(np.nanmean(wait, axis=(0, 1)) - np.nanmean(wait, axis=1).min(axis=0)).sum():.1f} us per call")
