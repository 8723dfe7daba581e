#!/usr/bin/env python3
"""Brute-force Hamming top-2 (BASELINE config 4): Q queries vs an M-row resident descriptor map on one GPU.
Reports batches/s, pair distances/s and the algorithmic GB/s (32*M + 48*Q bytes per batch)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_toolkit_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
m = api.Matcher(0)
rng = np.random.default_rng(1234)
db = rng.integers(0, 256, (args.rows, 32), dtype=np.uint8)
d = m.create_db(db)
for q in (1, 2, 4, 16, 64, 256, 2000):
    queries, _ = synth.knn_queries(db[:100000], q, seed=5678)
    dq = api.DeviceBuffer(q * 32).upload(queries)
    keys = api.DeviceBuffer(q * 16)
    m.knn2_dev(d, dq.ptr, q, keys.ptr)
    e0, e1 = api.Event(0), api.Event(0)
    m.set_async(True)                      # queue the batches back to back: launch latency stays off the clock
    e0.record(m)
    for _ in range(args.reps):
        m.knn2_dev(d, dq.ptr, q, keys.ptr)
    e1.record(m)
    m.wait()
    m.set_async(False)
    ms = e0.elapsed_ms(e1) / args.reps
    gb = (32 * args.rows + 48 * q) / 1e9
    print(f"Q={q:5d} M={args.rows}: {ms:8.3f} ms/batch  {gb / (ms / 1e3):8.1f} GB/s algorithmic  "
          f"{q * args.rows / (ms / 1e3) / 1e12:6.3f} T pair-distances/s")
