"""Host-side logic of the multi-GPU paths on CPU: world_size-2 `gloo` process groups.

Each rank computes its shard's candidates with the CPU oracle (standing in for the CUDA kernels, which need a
GPU), exchanges the packed keys through slam_toolkit_b200.sharding.gather_keys and merges them with numpy; the
result must equal the oracle run on the unsharded data.  What this pins: the block partition, the global index
bases, the key packing and the merge rules -- the parts that do not run on the GPU.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import oracle_c
    from slam_toolkit_b200 import sharding, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- brute-force kNN over a row-sharded map (odd size: uneven blocks) ------------------------------
        db = synth.knn_database(5001, seed=11)
        db[1700] = db[3400]            # an exact duplicate straddling the shard boundary: ties break on the global index
        queries, _ = synth.knn_queries(db, 257, seed=12)
        queries[0] = db[3400]
        start, stop = sharding.block(len(db), world, rank)
        local = oracle_c.knn2(queries, db[start:stop], idx_base=start)
        keys = torch.from_numpy(sharding.pack_knn_keys(local).view(np.int64))
        gathered = sharding.gather_keys(keys).numpy().view(np.uint64)           # world x q x 2
        merged = np.sort(gathered.transpose(1, 0, 2).reshape(len(queries), -1), axis=1)[:, :2]
        quad = np.empty((len(queries), 4), np.int32)
        for r in range(2):
            none = merged[:, r] == sharding.NO_KEY
            quad[:, 2 * r] = np.where(none, -1, (merged[:, r] & np.uint64(0xFFFFFFFF)).astype(np.int64))
            quad[:, 2 * r + 1] = np.where(none, 999999999, (merged[:, r] >> np.uint64(32)).astype(np.int64))
        full = oracle_c.knn2(queries, db)
        assert np.array_equal(quad, full), f"rank {rank}: sharded kNN differs on {(quad != full).any(axis=1).sum()} queries"
        assert quad[0, 0] == 1700 and quad[0, 1] == 0 and quad[0, 2] == 3400 and quad[0, 3] == 0

        # ---- ProjectionMatch over sharded map points ------------------------------------------------------
        rng = np.random.default_rng(5)
        m_kp = 300
        kps = np.zeros(m_kp, oracle_c.KP_DTYPE)
        kps["x"], kps["y"] = rng.uniform(20, 1220, m_kp), rng.uniform(20, 356, m_kp)
        kdesc = rng.integers(0, 256, (m_kp, 32), dtype=np.uint8)
        xy = np.stack([kps["x"], kps["y"]], 1)
        xw, mpd = synth.projection_scene(xy, kdesc, 4001, seed=6)
        xw[10], mpd[10] = xw[3000], mpd[3000]    # equal-distance conflict across shards: the later query must win
        cam = oracle_c.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0, 0, 0, 0], 1241, 376)
        rt = np.eye(3, 4)
        s0, s1 = sharding.block(len(xw), world, rank)
        to_q, dist_l = oracle_c.projection_match(xw[s0:s1], mpd[s0:s1], None, rt, cam, kps, kdesc, 50.0)
        gq = to_q.astype(np.int64) + s0
        k = np.where(to_q < 0, sharding.NO_KEY,
                     (dist_l.astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - gq.astype(np.uint64)))
        g = sharding.gather_keys(torch.from_numpy(k.view(np.int64))).numpy().view(np.uint64).min(axis=0)
        none = g == sharding.NO_KEY
        got_q = np.where(none, -1, (np.uint64(0xFFFFFFFF) - (g & np.uint64(0xFFFFFFFF))).astype(np.int64)).astype(np.int32)
        got_d = np.where(none, -1, (g >> np.uint64(32)).astype(np.int64)).astype(np.int32)
        ref_q, ref_d = oracle_c.projection_match(xw, mpd, None, rt, cam, kps, kdesc, 50.0)
        assert np.array_equal(got_q, ref_q) and np.array_equal(got_d, ref_d)
        assert (ref_q >= 0).sum() > 20

        # ---- frame sharding: blocks are contiguous, disjoint and cover the batch ---------------------------
        cover = torch.zeros(37, dtype=torch.int64)
        a, b = sharding.block(37, world, rank)
        cover[a:b] += 1
        dist.all_reduce(cover)
        assert cover.tolist() == [1] * 37
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_block_partition():
    from slam_toolkit_b200 import sharding
    for n in (0, 1, 7, 8, 4096, 10_000_000):
        for world in (1, 2, 3, 8):
            blocks = [sharding.block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert sharding.block(4096, 8, 3) == (1536, 2048)          # BASELINE config 3: 512 frames per GPU
    with pytest.raises(ValueError):
        sharding.block(5, 2, 2)


def test_world_size_2_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
