"""Multi-GPU parity as a pytest: spawns ``torch.distributed.run`` over every visible GPU (2, 4, 8) and runs
``tests/multigpu_check.py`` - sharded abc() / postpr() / SMC passes against the single-table oracle: accepted set
bit-exact, adjusted values 1e-9 - once over the NVLink mailboxes (exchanges inside the compute kernels, CUDA graph
replays) and once over the NCCL fallback (``ABCDLS_NO_PEER=1``: stage-by-stage tail, no graphs).  Skips on one GPU."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, env_extra, port, extra=()):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py"),
           *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith(("PASS", "FAIL"))]
    assert out.returncode == 0 and lines and all(l.startswith("PASS") for l in lines), \
        "\n".join(lines) + "\n" + out.stderr[-3000:]
    return lines


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["mailbox", "mailbox_fine", "nccl_fallback"])
def test_sharded_parity_over_all_visible_gpus(mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    # mailbox_fine: the second-stage counting sample (one more exchange: the window histograms) switched on at this
    # size (by default it starts at 2^24 rows per rank when there are several ranks)
    env = {"mailbox": {}, "mailbox_fine": {"ABCDLS_FINE_MIN": str(1 << 18)}, "nccl_fallback": {"ABCDLS_NO_PEER": "1"}}[mode]
    port = {"mailbox": 29541, "mailbox_fine": 29543, "nccl_fallback": 29542}[mode]
    lines = _run(world, env, port, ("--big", str(1 << 20)))
    assert len(lines) >= 8
    if mode != "nccl_fallback":
        assert any("graphs=" in l and not l.rstrip().endswith("graphs=0") for l in lines)
