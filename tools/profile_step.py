#!/usr/bin/env python3
"""One pass of every hot kernel for an ncu capture (profiles/r02_*): one resident 128-frame stereo sequence step (256
images per launch, the bench's shape), then the matchers at BASELINE sizes: brute-force top-2 with Q = 2000 / 1 / 2 / 4 against a
10 M-row map and ProjectionMatch of 500 k map points.  `--what step|match` picks one half; `--warm N` runs N passes.

    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_step python tools/profile_step.py --what step
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="step", choices=["step", "match"])
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--rows", type=int, default=10_000_000)
args = ap.parse_args()
W, H = synth.KITTI_W, synth.KITTI_H
cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, (0, 0, 0, 0), W, H)
if args.what == "step":
    F = args.frames
    Ls, Rs = [], []
    for s in range((F + 15) // 16):
        l, r = synth.stereo_sequence(s, min(16, F - 16 * s), 4)
        Ls.append(l); Rs.append(r)
    L, R = np.concatenate(Ls), np.concatenate(Rs)
    P = api.image_pitch(W)

    def pitched(a):
        o = np.zeros((a.shape[0], H, P), np.uint8)
        o[:, :, :W] = a
        return o
    ex = api.ORBextractor(max_images=2 * F)
    cap = ex.cap
    dl, dr = api.DeviceBuffer(F * P * H).upload(pitched(L)), api.DeviceBuffer(F * P * H).upload(pitched(R))
    spec = {"kps_l": 28 * cap * F, "desc_l": 32 * cap * F, "n_l": 4 * F, "kps_r": 28 * cap * F, "desc_r": 32 * cap * F, "n_r": 4 * F,
            "stereo_idx": 4 * cap * F, "stereo_dist": 4 * cap * F, "track_idx": 4 * cap * F, "track_dist": 4 * cap * F}
    bufs = {k: api.DeviceBuffer(v) for k, v in spec.items()}
    tp = api.TrackParams.make(cam, synth.KITTI_BASELINE, None, 50.0)
    for _ in range(args.warm):
        ex.stereo_sequence_dev(dl.ptr, dr.ptr, F, W, H, {k: b.ptr for k, b in bufs.items()}, tp, pitch=P)
    print("step done:", int(bufs["n_l"].download((F,), np.int32).sum()), "left keypoints")
else:
    m = api.Matcher(0)
    db = synth.knn_database(args.rows, seed=1234)
    q, _ = synth.knn_queries(db, 2000, seed=5678, hard_fraction=0.2)
    dbh = m.create_db(db)
    dq, keys = api.DeviceBuffer(q.nbytes).upload(q), api.DeviceBuffer(2000 * 16)
    for _ in range(args.warm):
        for nq in (2000, 1, 2, 4):
            m.knn2_dev(dbh, dq.ptr, nq, keys.ptr)
    ex = api.ORBextractor(max_images=2)
    kps, desc = ex.extract(synth.stereo_pair(0)[0])
    xy = np.stack([kps["x"], kps["y"]], 1)
    xw, mpd = synth.projection_scene(xy, desc, 500_000, seed=99)
    fr = api.Frame(m, kps, desc, cam)
    for _ in range(args.warm):
        got, _ = fr.ProjectionMatch(xw, mpd, None, np.array([0, 0, 0, 1, 0, 0, 0.0]), 50.0)
    print("match done:", int((got >= 0).sum()), "keypoints matched")
