import sys, torch, numpy as np
sys.path.insert(0, '.')
from abc_dls_b200 import synth
from abc_dls_b200.engine import AbcEngine
from oracle import abc_oracle as O
eng = AbcEngine(torch.device('cuda', 0))
par, ss, tgt = synth.abc_tables(1000, 3, 2, eng.device)
med, mad = eng.column_mad(ss, 1000)
x = ss.cpu().numpy()
print('med', med.cpu().numpy(), np.median(x, axis=1))
print('mad', mad.cpu().numpy(), [O.r_mad(x[j]) for j in range(3)])
r = eng._select(ss, 1000, None, 0, 999)
print('minmax', r.cpu().numpy(), x.min(1), x.max(1))
status = torch.zeros(1, dtype=torch.int32, device=eng.device)
dist = eng.k.new(1000)
eng.k.scale_dist(ss, 1000, 3, ss.stride(0), mad, tgt, dist, status)
print('status', status.item(), dist[:5].cpu().numpy())
_,_,_,od = O.scale_and_distance(tgt.cpu().numpy(), x.T)
print(od[:5])
