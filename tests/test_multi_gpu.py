"""N > 1 on real GPUs (NCCL): sharded kNN / ProjectionMatch / frame blocks against the oracle.
Needs >= 2 visible GPUs (`gpurun --gpus 2`); the single-GPU round-end run covers world == 1 below."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from slam_toolkit_b200 import api, sharding, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_classes_with_one_rank(oracle):
    """world == 1 (no process group): the sharded path degenerates to the plain one, through the same code."""
    if api.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    m = api.Matcher(0)
    db = synth.knn_database(50_000, seed=31)
    q, _ = synth.knn_queries(db, 100, seed=32)
    assert np.array_equal(sharding.ShardedDatabase(m, db, len(db)).knn2(q), oracle.knn2(q, db))


def test_two_ranks_nccl():
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mgpu ok" in r.stdout
