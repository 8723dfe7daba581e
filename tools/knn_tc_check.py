#!/usr/bin/env python3
"""Brute-force top-2 on the tensor cores (sfe_knn_tc.cu) against the XOR / POPC kernel and the CPU oracle, then timing.
Run under `timeout`: a mis-programmed mbarrier hand-off would otherwise spin until the kernel's own trap."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_c  # noqa: E402
from slam_toolkit_b200 import api, synth  # noqa: E402


def run(rows, q, seed, time_it=False):
    db = synth.knn_database(rows, seed=seed)
    qs, _ = synth.knn_queries(db, q, seed=seed + 1, hard_fraction=0.2)
    if rows > 300:
        db[rows - 1] = db[7]
        qs[1] = db[7]
    os.environ["SFE_KNN_TC"] = "1"
    m1 = api.Matcher(0)
    os.environ["SFE_KNN_TC"] = "0"
    m0 = api.Matcher(0)
    h1, h0 = m1.create_db(db), m0.create_db(db)
    a = m1.knn2(h1, qs)
    b = m0.knn2(h0, qs)
    ok = np.array_equal(a, b)
    msg = f"rows {rows} q {q}: tensor-core == popc kernel: {ok}"
    if rows * q <= 4e9:
        ref = oracle_c.knn2(qs, db, nthreads=os.cpu_count() or 8)
        msg += f", == oracle: {np.array_equal(a, ref)}"
        ok = ok and np.array_equal(a, ref)
    if not ok:
        bad = np.nonzero((a != b).any(1))[0]
        msg += f"  first differences at queries {bad[:5]}: {a[bad[:3]].tolist()} vs {b[bad[:3]].tolist()}"
    if time_it:
        dq, keys = api.DeviceBuffer(qs.nbytes).upload(qs), api.DeviceBuffer(q * 16)
        for m, h, name in ((m1, h1, "tensor-core"), (m0, h0, "popc")):
            m.knn2_dev(h, dq.ptr, q, keys.ptr)
            e0, e1 = api.Event(0), api.Event(0)
            m.set_async(True)
            e0.record(m)
            for _ in range(5):
                m.knn2_dev(h, dq.ptr, q, keys.ptr)
            e1.record(m)
            m.wait()
            m.set_async(False)
            msg += f"  {name} {e0.elapsed_ms(e1) / 5:.3f} ms"
    print(msg, flush=True)
    return ok


good = True
for rows, q in ((1000, 300), (128, 256), (50_001, 512), (50_001, 700), (200_003, 2000), (33, 1300)):
    good = run(rows, q, 11) and good
good = run(10_000_000, 2000, 1234, time_it=True) and good
good = run(1_250_000, 2000, 5, time_it=True) and good
sys.exit(0 if good else 1)
