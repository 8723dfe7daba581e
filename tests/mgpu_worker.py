"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), NCCL.  Every rank checks the sharded
results against the CPU oracle on the unsharded data and exits non-zero on any difference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import oracle_c
    from slam_toolkit_b200 import api, sharding, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = api.Matcher(local)
    # kNN over a row-sharded map, uneven blocks, duplicate rows across the boundary
    db = synth.knn_database(200_003, seed=21)
    db[70_000] = db[150_000]
    queries, _ = synth.knn_queries(db, 300, seed=22)
    queries[0] = db[150_000]
    a, b = sharding.block(len(db), world, rank)
    ref = oracle_c.knn2(queries, db, nthreads=4)
    for exchange in ("peer", "collective"):   # the library's NVLink peer-memory step, and NCCL all-gather as the carrier
        sdb = sharding.ShardedDatabase(m, db[a:b], len(db), exchange=exchange)
        for rep in range(3):                  # consecutive collectives reuse the two inbox parities
            got = sdb.knn2(queries[: 300 - 7 * rep])
            assert np.array_equal(got, ref[: 300 - 7 * rep]), f"rank {rank} {exchange}: kNN differs on {(got != ref[: 300 - 7 * rep]).any(axis=1).sum()} queries"
        if exchange == "peer":
            sdb.comm.status()
    # ProjectionMatch over sharded map points
    ex = api.ORBextractor(device=local, max_images=2)
    L, _ = synth.stereo_pair(0)
    kps, desc = ex.extract(L)
    xy = np.stack([kps["x"], kps["y"]], 1)
    xw, mpd = synth.projection_scene(xy, desc, 50_001, seed=23)
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    ocam = oracle_c.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    a, b = sharding.block(len(xw), world, rank)
    qt = np.array([0.003, -0.004, 0.001, 0.0, 0.05, -0.02, 0.2])
    qt[3] = np.sqrt(1 - (qt[:3] ** 2).sum())
    for pose in (np.array([0, 0, 0, 1, 0, 0, 0.0]), qt):
        rq, rd = oracle_c.projection_match(xw, mpd, None, pose, ocam, kps, desc, 50.0, grid=True)
        for exchange in ("peer", "collective"):
            lm = sharding.ShardedLocalMap(m, xw[a:b], mpd[a:b], len(xw), exchange=exchange)
            for rep in range(2):
                gq, gd = lm.projection_match(pose, cam, kps, desc, 50.0)
                assert np.array_equal(gq, rq) and np.array_equal(gd, rd), f"rank {rank} {exchange}: sharded projection match differs"
    assert (rq >= 0).sum() > 100
    # frame sharding: this rank's block of a batch, checked against the oracle
    seeds = list(range(world * 2 + 1))
    f0, f1 = sharding.block(len(seeds), world, rank)
    Ls = np.stack([synth.stereo_pair(s)[0] for s in seeds[f0:f1]])
    Rs = np.stack([synth.stereo_pair(s)[1] for s in seeds[f0:f1]])
    exb = api.ORBextractor(device=local, max_images=2 * len(Ls))
    out = exb.stereo_frames(Ls, Rs)
    o = oracle_c.Extractor()
    kl, dl = o.extract(Ls[0])
    assert out["n_l"][0] == len(kl) and np.array_equal(out["kps_l"][0, :len(kl)], kl) and np.array_equal(out["desc_l"][0, :len(kl)], dl)
    total = torch.tensor([int(out["n_l"].sum())], device=f"cuda:{local}")
    dist.all_reduce(total)
    dist.barrier()
    if rank == 0:
        print(f"mgpu ok: world {world}, {int(total.item())} left keypoints over {len(seeds)} frames")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
