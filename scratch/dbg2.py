import sys, torch, numpy as np
sys.path.insert(0, '.')
from abc_dls_b200.engine import AbcEngine
eng = AbcEngine(torch.device('cuda', 0))
k = eng.k
x = torch.arange(1000, dtype=torch.float64, device='cuda')[None, :].contiguous() + 1.0
ws = k.select_workspace(1)
k.select_begin(ws, 1, x, 1000, 1000, None, None, 499, 500)
torch.cuda.synchronize()
job = ws[:88].cpu().view(torch.int64)
print('job after begin', job.tolist())
hist = k.select_hist_tensor(ws, 1)
print('hist numel', hist.numel(), 'offset', hist.data_ptr() - ws.data_ptr())
k.select_hist(ws, 1, 1000, 0)
torch.cuda.synchronize()
h = hist.cpu().view(2, 2048)
print('hist sums', h.sum(1).tolist(), 'nonzero bins', h[0].nonzero().view(-1).tolist()[:10])
k.select_pick(ws, 1, 0)
torch.cuda.synchronize()
print('job after pick', ws[:88].cpu().view(torch.int64).tolist())
for p in range(1, 6):
    k.select_hist(ws, 1, 1000, p); torch.cuda.synchronize()
    h = hist.cpu().view(2, 2048)
    print(p, 'hist sums', h.sum(1).tolist())
    k.select_pick(ws, 1, p); torch.cuda.synchronize()
    print(p, 'job', ws[:88].cpu().view(torch.int64).tolist())
out = k.new(1, 2)
k.select_end(ws, 1, out); torch.cuda.synchronize()
print(out)
