"""Small collectives used between kernel stages.  world == 1 => every call is a no-op.

One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU
tests of the host logic).  Messages on this path are tiny (histograms, counts, a (d+1) x (d+1+P) fp64
partial, 2P min/max) so every call is latency bound; callers pack what they can into one tensor.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


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
        # set by the engine while it records a call as CUDA-graph segments: collectives are not captured, they end the
        # current segment, run eagerly (and are replayed eagerly) between two graph launches
        self.recorder = None

    def _do(self, fn):
        self.calls += 1
        if self.recorder is not None:
            self.recorder.split(fn)
        else:
            fn()

    def _red(self, t: torch.Tensor, op):
        if self.world > 1:
            self._do(lambda: dist.all_reduce(t, op=op, group=self.group))
        return t

    def allreduce_sum(self, t):
        return self._red(t, dist.ReduceOp.SUM)

    def allreduce_min(self, t):
        return self._red(t, dist.ReduceOp.MIN)

    def allreduce_max(self, t):
        return self._red(t, dist.ReduceOp.MAX)

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        """(world, *t.shape) stacked copies, rank order."""
        if self.world == 1:
            return t[None]
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        src = t.contiguous().view(-1)
        self._do(lambda: dist.all_gather_into_tensor(out.view(-1), src, group=self.group))
        return out

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
