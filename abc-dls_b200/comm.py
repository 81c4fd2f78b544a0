"""Small collectives used between kernel stages.  world == 1 => every call is a no-op.

One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU
tests of the host logic).  Messages on this path are tiny (histograms, counts, a (d+1) x (d+1+P) fp64
partial, 2P min/max) so every call is latency bound; callers pack what they can into one tensor.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List

import torch
import torch.distributed as dist

PEER_SLOT_BYTES = 2 << 20   # largest message that goes through the NVLink mailboxes (bigger ones use NCCL)


class Comm:
    def __init__(self, group=None, device=None):
        self.group = group
        self.device = device
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.calls = 0
        # set by the engine while it records a call as CUDA-graph segments: NCCL collectives are not captured, they end
        # the current segment, run eagerly (and are replayed eagerly) between two graph launches
        self.recorder = None
        # NVLink peer-memory mailboxes (csrc/peer.cu): every small exchange is ONE kernel, no NCCL, graph capturable
        self.peer = None
        self.peer_calls = 0
        self.fused_calls = 0   # exchanges that ran INSIDE a compute kernel's last block (csrc/tail.cu)
        if (self.world > 1 and device is not None and torch.device(device).type == "cuda"
                and os.environ.get("ABCDLS_NO_PEER", "0") != "1"):
            self._init_peer()

    def _init_peer(self):
        from . import _capi
        lib = _capi.load()
        hb = lib.abcdls_peer_handle_bytes()
        handle = (C.c_ubyte * hb)()
        p = C.c_void_p()
        ok = lib.abcdls_peer_create(C.byref(p), self.rank, self.world, PEER_SLOT_BYTES, handle) == 0
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
        allh = torch.empty(self.world * hb, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)          # bootstrap only: the IPC handles
        blob = bytes(allh.cpu().tolist())
        if ok:
            ok = lib.abcdls_peer_connect(p, blob) == 0
        bad = torch.tensor([0 if ok else 1], dtype=torch.int64, device=self.device)
        dist.all_reduce(bad, group=self.group)                             # every rank uses the mailboxes, or none
        if int(bad.item()) == 0:
            self.peer, self._lib = p, lib
        elif p:
            lib.abcdls_peer_destroy(p)
        dist.barrier(group=self.group)

    def trace(self):
        """(records (n, 4) uint64 ns: entry / published / flags seen / exit of the last exchanges, oldest first) when
        the mailboxes were created under ABCDLS_PEER_TRACE=1, else None (profiles/tools/peer_timeline.py)."""
        if self.peer is None:
            return None
        import numpy as np
        buf = np.zeros((4096, 4), dtype=np.uint64)
        n = C.c_uint64(0)
        cap = self._lib.abcdls_peer_trace_read(self.peer, buf.ctypes.data_as(C.c_void_p), C.byref(n))
        if cap == 0:
            return None
        k = int(n.value)
        idx = [(i % cap) for i in range(max(1, k - cap + 1), k + 1)]
        return buf[idx]

    def close(self):
        if self.peer is not None:
            self._lib.abcdls_peer_destroy(self.peer)
            self.peer = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _do(self, fn):
        self.calls += 1
        if self.recorder is not None:
            self.recorder.split(fn)
        else:
            fn()

    def _red(self, t: torch.Tensor, op, opcode):
        if self.world > 1:
            if (self.peer is not None and t.is_cuda and t.is_contiguous() and t.dtype in (torch.float64, torch.int64)
                    and t.numel() * 8 <= PEER_SLOT_BYTES):
                rc = self._lib.abcdls_peer_allreduce(self.peer, C.c_void_p(t.data_ptr()), t.numel(),
                                                     0 if t.dtype == torch.float64 else 1, opcode, self._stream())
                if rc != 0:
                    raise RuntimeError("abcdls_peer_allreduce failed")
                self.calls += 1
                self.peer_calls += 1
            else:
                self._do(lambda: dist.all_reduce(t, op=op, group=self.group))
        return t

    def allreduce_sum(self, t):
        return self._red(t, dist.ReduceOp.SUM, 0)

    def allreduce_min(self, t):
        return self._red(t, dist.ReduceOp.MIN, 1)

    def allreduce_max(self, t):
        return self._red(t, dist.ReduceOp.MAX, 2)

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        """(world, *t.shape) stacked copies, rank order."""
        if self.world == 1:
            return t[None]
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        src = t.contiguous().view(-1)
        nbytes = src.numel() * src.element_size()
        if self.peer is not None and t.is_cuda and nbytes % 8 == 0 and 0 < nbytes <= PEER_SLOT_BYTES:
            rc = self._lib.abcdls_peer_allgather(self.peer, C.c_void_p(src.data_ptr()), nbytes,
                                                 C.c_void_p(out.data_ptr()), self._stream())
            if rc != 0:
                raise RuntimeError("abcdls_peer_allgather failed")
            self.calls += 1
            self.peer_calls += 1
            if src.data_ptr() != t.data_ptr():
                out._abcdls_keep = src      # a temporary contiguous copy must outlive the enqueued kernel
        else:
            self._do(lambda: dist.all_gather_into_tensor(out.view(-1), src, group=self.group))
        return out

    def allgather_ragged(self, vals: torch.Tensor, cnt: torch.Tensor):
        """vals (nj * cap,) float64 = nj rows of which only the first cnt[j] (int32) entries exist ->
        ((world, nj * cap) values, (world, nj) counts).  Over the mailboxes only the existing entries cross NVLink and
        both arrays travel in ONE exchange; otherwise two plain all-gathers."""
        nj = cnt.numel()
        cap = vals.numel() // nj
        if (self.world > 1 and self.peer is not None and vals.is_cuda and vals.is_contiguous() and cnt.is_contiguous()
                and (nj + nj * cap) * 8 <= PEER_SLOT_BYTES):
            gv = torch.empty((self.world, nj * cap), dtype=torch.float64, device=vals.device)
            gc = torch.empty((self.world, nj), dtype=torch.int32, device=vals.device)
            rc = self._lib.abcdls_peer_allgather_ragged(self.peer, C.c_void_p(vals.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                                        nj, cap, C.c_void_p(gv.data_ptr()), C.c_void_p(gc.data_ptr()),
                                                        self._stream())
            if rc != 0:
                raise RuntimeError("abcdls_peer_allgather_ragged failed")
            self.calls += 1
            self.peer_calls += 1
            return gv, gc
        return self.allgather(vals).contiguous(), self.allgather(cnt).contiguous()

    def allgather_ints(self, vals: List[int]) -> List[List[int]]:
        """Host integers from every rank (shapes / row counts); one small collective + a sync."""
        if self.world == 1:
            return [list(vals)]
        t = torch.tensor(vals, dtype=torch.int64, device=self.device)
        return self.allgather(t).cpu().tolist()

    def allreduce_or_int(self, v: int) -> int:
        if self.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.int64, device=self.device)
        # bitwise OR over ranks == max per bit; status words are small bit sets
        bits = torch.tensor([(v >> b) & 1 for b in range(16)], dtype=torch.int64, device=self.device)
        self.allreduce_max(bits)
        del t
        return int(sum(int(x) << b for b, x in enumerate(bits.cpu().tolist())))

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)
