"""Data-parallel path on real GPUs (needs >= 2 visible devices; skipped otherwise): toist_b200.util.dist's drop-in for
DistributedDataParallel (reference main.py:336) must leave every rank with bit-identical gradients equal to the mean of
the ranks' local gradients, in eager mode and under CUDA-graph replay.  The check itself is tools/ddp_check.py, launched
with torchrun like the benchmark."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("graphs", ["1"])  # the CUDA-graph path is what bench.py runs; DDP_CHECK_GRAPHS=0 checks eager mode by hand
def test_flat_gradient_all_reduce_two_ranks(graphs):
    env = dict(os.environ, DDP_CHECK_GRAPHS=graphs, NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tools" / "ddp_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "gradients identical across ranks" in res.stdout
