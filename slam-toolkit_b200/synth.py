"""Seeded synthetic KITTI-shaped inputs.  numpy only, frozen.

KITTI itself is not available offline; the reference's workload shape is
1241x376 8-bit gray stereo (reference src/dataset.cpp:103-104).  Every generator
here is a pure function of its seed so CPU oracle, CUDA path and golden fixtures
provably see identical bytes (tests/golden/synth_sha256.json pins seeds 0-7).
"""
from __future__ import annotations

import numpy as np

KITTI_W, KITTI_H = 1241, 376
DISPARITY = 24
_EXTRA = 128
# KITTI seq-00 intrinsics / baseline as hard-coded by reference src/dataset.cpp:87-101
KITTI_FX = 7.188560000000e+02
KITTI_FY = 7.188560000000e+02
KITTI_CX = 6.071928e+02
KITTI_CY = 1.852157000000e+02
KITTI_BASELINE = 3.861448000000e+02 / 7.188560000000e+02


def _blur_sigma08(a: np.ndarray) -> np.ndarray:
    """Separable Gaussian, sigma 0.8, radius 3, edge-replicated (float32)."""
    r = 3
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2 * 0.8 * 0.8))
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(a, ((0, 0), (r, r)), mode="edge")
    h = np.zeros_like(a)
    for i in range(2 * r + 1):
        h += k[i] * p[:, i:i + a.shape[1]]
    p = np.pad(h, ((r, r), (0, 0)), mode="edge")
    v = np.zeros_like(a)
    for i in range(2 * r + 1):
        v += k[i] * p[i:i + a.shape[0], :]
    return v


def canvas(seed: int, width: int = KITTI_W, height: int = KITTI_H) -> np.ndarray:
    """u8 canvas height x (width+128); stereo views are column windows of it."""
    rng = np.random.default_rng(seed)
    cw = width + _EXTRA
    xs = np.arange(cw, dtype=np.float32)[None, :]
    ys = np.arange(height, dtype=np.float32)[:, None]
    c = (96.0 + 48.0 * np.sin(xs / 97.0) + 32.0 * np.cos(ys / 53.0)).astype(np.float32)
    n_rect = max(1, int(round(900 * (cw * height) / float((KITTI_W + _EXTRA) * KITTI_H))))
    for _ in range(n_rect):
        x = int(rng.integers(0, cw - 8))
        y = int(rng.integers(0, max(1, height - 8)))
        w = int(rng.integers(6, 70))
        h = int(rng.integers(6, 50))
        val = float(rng.integers(0, 256))
        c[y:y + h, x:x + w] = val
    c = _blur_sigma08(c)
    c = c + rng.normal(0.0, 2.0, c.shape).astype(np.float32)
    return np.clip(np.rint(c), 0, 255).astype(np.uint8)


def stereo_pair(seed: int, width: int = KITTI_W, height: int = KITTI_H):
    """(left, right) contiguous u8 images; right = left shifted by 24 px
    (uniform disparity so every true match has dy=0, 0<=dx<=100; cf. the
    filter at reference src/matcher.cpp:103-110)."""
    c = canvas(seed, width, height)
    left = np.ascontiguousarray(c[:, :width])
    right = np.ascontiguousarray(c[:, DISPARITY:DISPARITY + width])
    return left, right


def knn_database(m: int, seed: int = 1234) -> np.ndarray:
    """m x 32 u8 random descriptors (BASELINE config 4)."""
    return np.random.default_rng(seed).integers(0, 256, (m, 32), dtype=np.uint8)


def knn_queries(db: np.ndarray, q: int, seed: int = 5678, max_flips: int = 40, hard_fraction: float = 0.0,
                hard_flips=(45, 90)):
    """q rows picked from db with 0..max_flips random bit flips each.  Among 10 M random 256-bit rows the second-best
    distance of such a query is ~85, so up to 40 flips always pass the ratio test of reference src/matcher.cpp:125
    (2 * dist0 < dist1); `hard_fraction` of the queries get hard_flips[0]..hard_flips[1] flips instead, so that the
    test both passes and fails."""
    rng = np.random.default_rng(seed)
    rows = rng.choice(db.shape[0], q, replace=db.shape[0] < q)
    out = db[rows].copy()
    for i in range(q):
        nf = int(rng.integers(0, max_flips + 1))
        bits = rng.integers(0, 256, nf)
        for b in bits:
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    if hard_fraction > 0:
        rng2 = np.random.default_rng(seed + 1)
        for i in np.nonzero(rng2.random(q) < hard_fraction)[0]:
            out[i] = db[rows[i]]
            for b in rng2.choice(256, int(rng2.integers(hard_flips[0], hard_flips[1] + 1)), replace=False):
                out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out, rows


def stereo_sequence(seed: int, n: int, step: int = 4, width: int = KITTI_W, height: int = KITTI_H):
    """`n` consecutive stereo frames of scene `seed`: the camera slides `step` px per frame along the canvas
    (left_k = columns [k step, k step + W), right_k = the same window 24 px further), so consecutive frames share their
    content and the tracker's ProjectionMatch has true correspondences within a few pixels.  n * step <= 104."""
    assert (n - 1) * step + DISPARITY <= _EXTRA, "sequence runs off the canvas"
    c = canvas(seed, width, height)
    left = np.stack([c[:, k * step:k * step + width] for k in range(n)])
    right = np.stack([c[:, k * step + DISPARITY:k * step + DISPARITY + width] for k in range(n)])
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


def projection_scene(kps_xy: np.ndarray, desc: np.ndarray, n_points: int, seed: int = 99,
                     width: int = KITTI_W, height: int = KITTI_H):
    """Map points for BASELINE config 5: Xw uniform in the KITTI frustum
    (z~U[2,80] m, back-projected from uniform pixels), descriptors = 10 % noisy
    copies of frame descriptors placed within 20 px of that keypoint, 90 % random.
    Returns (Xw float64 n x 3, desc u8 n x 32).  Pose is identity."""
    rng = np.random.default_rng(seed)
    n_kp = kps_xy.shape[0]
    u = rng.uniform(0, width, n_points)
    v = rng.uniform(0, height, n_points)
    z = rng.uniform(2.0, 80.0, n_points)
    d = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
    if n_kp > 0:
        n_copy = n_points // 10
        sel = rng.choice(n_points, n_copy, replace=False)
        src = rng.integers(0, n_kp, n_copy)
        ang = rng.uniform(0, 2 * np.pi, n_copy)
        rad = rng.uniform(0, 20.0, n_copy)
        u[sel] = kps_xy[src, 0] + rad * np.cos(ang)
        v[sel] = kps_xy[src, 1] + rad * np.sin(ang)
        d[sel] = desc[src]
        nflip = rng.integers(0, 31, n_copy)
        for i in range(n_copy):
            for b in rng.integers(0, 256, int(nflip[i])):
                d[sel[i], b >> 3] ^= np.uint8(1 << (b & 7))
    X = np.empty((n_points, 3), dtype=np.float64)
    X[:, 0] = (u - KITTI_CX) / KITTI_FX * z
    X[:, 1] = (v - KITTI_CY) / KITTI_FY * z
    X[:, 2] = z
    return X, np.ascontiguousarray(d)


def vocabulary(k: int = 10, L: int = 4, seed: int = 7, stop_fraction: float = 0.05, early_leaf: float = 0.03):
    """A synthetic DBoW2-style vocabulary tree (the real ORBvoc.txt is not redistributable with the reference):
    branching k, depth L, breadth-first node numbering as the k-means builder produces, random descriptors that share
    most bits with their parent (so descents are decided by few bits and ties occur), random idf weights with a few
    stopped words (weight 0) and a few branches that end above depth L.
    Returns (parent int32[n], is_leaf u8[n], desc u8[n,32], weight f64[n], L)."""
    rng = np.random.default_rng(seed)
    parent, is_leaf, desc, weight, depth = [0], [0], [np.zeros(32, np.uint8)], [0.0], [0]
    frontier = [0]
    for level in range(1, L + 1):
        nxt = []
        for p in frontier:
            for _ in range(k):
                d = desc[p].copy()
                for b in rng.integers(0, 256, 24 if level > 1 else 128):
                    d[b >> 3] ^= np.uint8(1 << (b & 7))
                leaf = level == L or rng.uniform() < early_leaf
                parent.append(p)
                is_leaf.append(1 if leaf else 0)
                desc.append(d)
                weight.append(0.0 if (leaf and rng.uniform() < stop_fraction) else (float(rng.uniform(0.1, 9.0)) if leaf else 0.0))
                depth.append(level)
                if not leaf:
                    nxt.append(len(parent) - 1)
        frontier = nxt
    return (np.array(parent, np.int32), np.array(is_leaf, np.uint8), np.stack(desc).astype(np.uint8),
            np.array(weight, np.float64), L)
