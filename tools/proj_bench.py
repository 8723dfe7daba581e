#!/usr/bin/env python3
"""ProjectionMatch (BASELINE config 5): N map points against the keypoints of one synthetic frame, r = 50 px, resident."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=500_000)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--radius", type=float, default=50.0)
args = ap.parse_args()
L, _ = synth.stereo_pair(0)
ex = api.ORBextractor(2000, 1.2, 8, 20, 7, device=0, max_images=2)
kps, desc = ex.extract(L)
xy = np.stack([kps["x"], kps["y"]], axis=1).astype(np.float64)
xw, mpd = synth.projection_scene(xy, desc, args.points, seed=99)
cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, (0, 0, 0, 0), synth.KITTI_W, synth.KITTI_H)
m = api.Matcher(0)
n, mk = args.points, len(kps)
d_xw = api.DeviceBuffer(n * 24).upload(xw)
d_mpd = api.DeviceBuffer(n * 32).upload(mpd)
d_kps = api.DeviceBuffer(mk * 28).upload(np.ascontiguousarray(kps))
d_kd = api.DeviceBuffer(mk * 32).upload(np.ascontiguousarray(desc))
d_q, d_d = api.DeviceBuffer(mk * 4), api.DeviceBuffer(mk * 4)


def call():
    m.projection_match_dev(d_xw.ptr, d_mpd.ptr, None, n, np.eye(4), cam, d_kps.ptr, d_kd.ptr, mk, args.radius, d_q.ptr, d_d.ptr)


call()
e0, e1 = api.Event(0), api.Event(0)
m.set_async(True)
e0.record(m)
for _ in range(args.reps):
    call()
e1.record(m)
m.wait()
m.set_async(False)
ms = e0.elapsed_ms(e1) / args.reps
alg = n * 56 + mk * 48
print(f"N={n} M={mk} r={args.radius}: {ms:.4f} ms/call  {n / ms / 1e6:.2f} G points/s  {alg / ms / 1e6:.1f} GB/s algorithmic  "
      f"matched {(d_q.download((mk,), np.int32) >= 0).sum()}")
