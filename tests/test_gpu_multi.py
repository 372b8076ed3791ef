"""Multi-GPU checks that need real peers (skipped on a single-GPU box): run as torchrun subprocesses."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, n, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_bn_exchange_equals_allreduce():
    r = _torchrun("tests/multi/peer_bn_exchange.py", 2)
    assert r.returncode == 0 and "PEER_BN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
