"""Multi-GPU checks that need real peers (skipped on a single-GPU box): run as torchrun subprocesses."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, n, timeout=240, env=None, port=29533):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env={**os.environ, **(env or {})})


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_bn_exchange_equals_allreduce():
    r = _torchrun("tests/multi/peer_bn_exchange.py", 2)
    assert r.returncode == 0 and "PEER_BN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("n", [2, 4])
def test_sharded_training_step_equals_global_batch(n):
    """N-rank batch-sharded MM-mode step (sync-BN over peer memory, averaged gradient bucket) == the global batch on one
    rank: loss, predictions, every gradient, BatchNorm buffers (tests/multi/sharded_step_parity.py)."""
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = _torchrun("tests/multi/sharded_step_parity.py", n, timeout=600, port=29541 + n)
    assert r.returncode == 0 and "SHARDED_STEP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-4000:]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_nccl_halo_exchange_equals_whole_graph(n):
    """Graph-partitioned processor over N real ranks (dist.HaloExchange over NCCL) == the whole graph, >= 100 k nodes
    (tests/multi/halo_parity.py)."""
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    r = _torchrun("tests/multi/halo_parity.py", n, timeout=600, port=29551 + n)
    assert r.returncode == 0 and "HALO_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-4000:]
