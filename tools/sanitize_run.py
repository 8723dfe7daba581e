#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, both the TMA and the
plain-load variants, tiny inputs so the instrumented run stays short."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth  # noqa: E402

for no_tma in ("0", "1"):
    os.environ["SFE_NO_TMA"] = no_tma
    for (w, h, nf, nl) in ((320, 240, 500, 4), (161, 131, 300, 3), (1241, 376, 2000, 8)):
        L, R = synth.stereo_pair(7, w, h)
        ex = api.ORBextractor(nf, 1.2, nl, 20, 7, max_images=4)
        out = ex.stereo_frames(np.stack([L, R]), np.stack([R, L]))
        if w == 320:  # the one-pair call: eager, graph capture, replay; pinned arrays (stored by the copy kernel) and pageable ones (staged)
            for pinned in (True, False):
                res = ex.alloc_stereo_out(1, pinned=pinned)
                for _ in range(3):
                    one = ex.stereo_frames(L[None], R[None], res)
                assert int(one["n_l"][0]) == int(out["n_l"][0]), (one["n_l"], out["n_l"])
        k, d = ex.extract(L)
        print(no_tma, w, h, int(out["n_l"].sum()), len(k), flush=True)
m = api.Matcher(0)
db = synth.knn_database(5000, seed=1)
q, _ = synth.knn_queries(db, 300, seed=2)
dbh = m.create_db(db)
print("knn", m.knn2(dbh, q)[:2].tolist(), m.knn2(dbh, q[:3])[:1].tolist(), flush=True)
xy = np.stack([k["x"], k["y"]], 1)
xw, mpd = synth.projection_scene(xy, d, 3000, seed=3)
cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [-0.05, 0.01, 0.001, -0.002], 1241, 376)
f = api.Frame(m, k, d, cam)
print("proj", int((f.ProjectionMatch(xw, mpd, None, np.eye(3, 4), 50.0)[0] >= 0).sum()), f.normalized()[:1].tolist(),
      len(f.SearchRadius([[600.0, 180.0]], 50.0)[0]), f.SearchNeareast([[600.0, 180.0]])[0].tolist(), flush=True)
