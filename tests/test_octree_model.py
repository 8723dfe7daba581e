"""The array/prefix-sum reformulation of DistributeOctTree that the CUDA kernel
transcribes (tools/octree_model.py) against the literal std::list emulation in
the C oracle.  CPU only."""
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import octree_model
from slam_toolkit_b200 import synth


def _run(oracle, xyr, W, H, want, w_cell, h_cell, shuffle_seed=0):
    ref = oracle.distribute(xyr, 16, 16 + W, 16, 16 + H, want)
    # the model must be insensitive to candidate order: feed it a shuffled list
    order = np.random.default_rng(shuffle_seed).permutation(len(xyr))
    s = xyr[order]
    kept = octree_model.distribute(s[:, 0].astype(int), s[:, 1].astype(int), s[:, 2].astype(int), W, H, want, w_cell, h_cell)
    assert np.array_equal(s[kept].reshape(-1, 3), ref)


def test_model_on_real_candidates(oracle):
    ex = oracle.Extractor()
    L, _ = synth.stereo_pair(2)
    ex.extract(L)
    per = ex.tables()["per_level"]
    for l in range(8):
        lw, lh = ex.level_size(1241, 376, l)
        W, H = lw - 32, lh - 32
        w_cell, h_cell = int(np.ceil(W / (W // 30))), int(np.ceil(H / (H // 30)))
        _run(oracle, ex.candidates(l), W, H, int(per[l]), w_cell, h_cell, l)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 500), st.integers(1, 400), st.integers(60, 500), st.integers(31, 140), st.integers(0, 2**31))
def test_model_random(oracle, n, want, W, H, seed):
    if round(W / H) < 1:
        H = W
    rng = np.random.default_rng(seed)
    w_cell, h_cell = int(np.ceil(W / (W // 30))), int(np.ceil(H / (H // 30)))
    pts = set()
    n = min(n, (W - 6) * (H - 6) // 2)
    while len(pts) < n:
        pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
    # few distinct responses -> many max-response ties inside nodes
    key = lambda p: octree_model.order_key(p[0], p[1], w_cell, h_cell)
    xyr = np.array([(x, y, int(rng.integers(7, 12))) for x, y in sorted(pts, key=key)], np.float32)
    _run(oracle, xyr, W, H, want, w_cell, h_cell, seed)
