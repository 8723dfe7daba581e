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
    for exchange in ("peer", "collective"):
        assert np.array_equal(sharding.ShardedDatabase(m, db, len(db), exchange=exchange).knn2(q), oracle.knn2(q, db))


def test_one_process_driving_two_gpus(oracle):
    """sfe_comm_create_local: a single host thread enqueues the sharded kNN and ProjectionMatch on both GPUs (the calls
    return once enqueued), then waits; results equal the oracle on the unsharded data, on both ranks."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    comms = api.Comm.create_local([0, 1])
    ms = [api.Matcher(0), api.Matcher(1)]
    db = synth.knn_database(120_001, seed=41)
    q, _ = synth.knn_queries(db, 257, seed=42)
    ref = oracle.knn2(q, db, nthreads=4)
    bounds = [sharding.block(len(db), 2, r) for r in range(2)]
    shards = [ms[r].create_db(db[a:b], idx_base=a) for r, (a, b) in enumerate(bounds)]
    d_q = [api.DeviceBuffer(q.nbytes, r).upload(q) for r in range(2)]
    d_o = [api.DeviceBuffer(len(q) * 16, r) for r in range(2)]
    for rep in range(3):
        for r in range(2):
            ms[r].knn2_sharded(comms[r], shards[r], d_q[r].ptr, len(q), d_o[r].ptr)
        for r in range(2):
            ms[r].wait()
            assert np.array_equal(d_o[r].download((len(q), 4), np.int32), ref), (rep, r)
    ex = api.ORBextractor(max_images=2)
    kps, desc = ex.extract(synth.stereo_pair(0)[0])
    xy = np.stack([kps["x"], kps["y"]], 1)
    xw, mpd = synth.projection_scene(xy, desc, 40_001, seed=43)
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    ocam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    pose = np.array([0, 0, 0, 1, 0, 0, 0.0])
    rq, rd = oracle.projection_match(xw, mpd, None, pose, ocam, kps, desc, 50.0, grid=True)
    frames = [api.Frame(ms[r], kps, desc, cam) for r in range(2)]
    pb = [sharding.block(len(xw), 2, r) for r in range(2)]
    d_x = [api.DeviceBuffer((b - a) * 24, r).upload(np.ascontiguousarray(xw[a:b])) for r, (a, b) in enumerate(pb)]
    d_d = [api.DeviceBuffer((b - a) * 32, r).upload(np.ascontiguousarray(mpd[a:b])) for r, (a, b) in enumerate(pb)]
    d_tq = [api.DeviceBuffer(len(kps) * 4, r) for r in range(2)]
    d_td = [api.DeviceBuffer(len(kps) * 4, r) for r in range(2)]
    for r, (a, b) in enumerate(pb):
        ms[r].projection_match_sharded(comms[r], frames[r], d_x[r].ptr, d_d[r].ptr, None, b - a, a, pose, 50.0, d_tq[r].ptr, d_td[r].ptr)
    for r in range(2):
        ms[r].wait()
        assert np.array_equal(d_tq[r].download((len(kps),), np.int32), rq) and np.array_equal(d_td[r].download((len(kps),), np.int32), rd)
        comms[r].status()


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
