"""Seeded synthetic simulation tables: identical at 1/2/4/8 GPUs.

Rows come in fixed chunks of 2**20; chunk c is drawn from ``torch.Generator(device).manual_seed(SEED + c)``
so a rank can materialise exactly the row block it owns.  Values are rounded through float32 and stored
as float64 (Keras float32 predictions widened by pandas / R, reference ABC.py:1849, 498).  Parameter
ranges cycle through the 24 rows of the reference's ``examples/hardrange.csv``.
"""
from __future__ import annotations

from typing import Tuple

import torch

SEED = 20240919
CHUNK = 1 << 20

# (min, max) of examples/hardrange.csv, rows N_A .. XMix
HARDRANGE = [
    (5000.0, 25000.0), (10000.0, 150000.0), (10000.0, 150000.0), (10000.0, 150000.0), (500.0, 5000.0),
    (500.0, 5000.0), (500.0, 40000.0), (500.0, 5000.0), (500.0, 40000.0), (10.0, 50.0), (5.0, 30.0), (5.0, 40.0),
    (5.0, 120.0), (5.0, 50.0), (5.0, 50.0), (5.0, 220.0), (5.0, 700.0), (330.0, 450.0), (120.0, 250.0),
    (450.0, 700.0), (0.05, 0.95), (0.01, 0.03), (0.0, 0.02), (0.0, 0.1),
]


def param_ranges(P: int, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    lo = torch.tensor([HARDRANGE[p % 24][0] for p in range(P)], dtype=torch.float64, device=device)
    hi = torch.tensor([HARDRANGE[p % 24][1] for p in range(P)], dtype=torch.float64, device=device)
    return lo, hi


def _f32(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float32).to(torch.float64)


def abc_tables(n_rows: int, d: int, P: int, device, row_start: int = 0, seed: int = SEED, lo=None, hi=None):
    """(param (P, n_rows), sumstat (d, n_rows), target (d,)) column-major float64 tables for global rows
    [row_start, row_start + n_rows).  ``lo``/``hi`` override the prior box (SMC rounds)."""
    device = torch.device(device)
    plo, phi = param_ranges(P, device)
    if lo is not None:
        plo, phi = lo.to(device=device, dtype=torch.float64), hi.to(device=device, dtype=torch.float64)
    span = phi - plo
    param = torch.empty((P, n_rows), dtype=torch.float64, device=device)
    ss = torch.empty((d, n_rows), dtype=torch.float64, device=device)
    jmap = torch.arange(d, device=device) % P
    c0, c1 = row_start // CHUNK, (row_start + n_rows - 1) // CHUNK
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed + c)
        u = torch.rand((P, CHUNK), generator=g, dtype=torch.float64, device=device)
        z = torch.randn((d, CHUNK), generator=g, dtype=torch.float64, device=device)
        theta = _f32(plo[:, None] + span[:, None] * u)
        s = _f32(theta[jmap] + 0.1 * span[jmap][:, None] * z)
        a = max(row_start, c * CHUNK)
        b = min(row_start + n_rows, (c + 1) * CHUNK)
        param[:, a - row_start: b - row_start] = theta[:, a - c * CHUNK: b - c * CHUNK]
        ss[:, a - row_start: b - row_start] = s[:, a - c * CHUNK: b - c * CHUNK]
        del u, z, theta, s
    target = _f32(plo[jmap] + 0.37 * span[jmap])
    return param, ss, target


def postpr_tables(n_rows: int, device, K: int = 3, row_start: int = 0, seed: int = SEED):
    """(model_id int32 (n_rows,), sumstat (K, n_rows), target (K,)): K-model softmax rounded to 5 decimals
    (reference ABC.py:792-793) — exact duplicate rows and distance ties are common by construction."""
    device = torch.device(device)
    ss = torch.empty((K, n_rows), dtype=torch.float64, device=device)
    c0, c1 = row_start // CHUNK, (row_start + n_rows - 1) // CHUNK
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed + 7919 + c)
        rows = torch.arange(c * CHUNK, (c + 1) * CHUNK, device=device)
        z = torch.randn((K, CHUNK), generator=g, dtype=torch.float32, device=device)
        logits = z + 2.0 * (torch.arange(K, device=device)[:, None] == (rows % K)[None, :]).to(torch.float32)
        sm = torch.round(torch.softmax(logits, dim=0), decimals=5).to(torch.float64)
        a = max(row_start, c * CHUNK)
        b = min(row_start + n_rows, (c + 1) * CHUNK)
        ss[:, a - row_start: b - row_start] = sm[:, a - c * CHUNK: b - c * CHUNK]
    ids = ((torch.arange(row_start, row_start + n_rows, device=device)) % K).to(torch.int32)
    tgt = torch.tensor([0.2, 0.5, 0.3] + [0.0] * max(0, K - 3), dtype=torch.float64, device=device)[:K]
    return ids, ss, tgt
