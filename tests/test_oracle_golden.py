"""The C oracle (oracle/orb_oracle.c) against the committed fixtures that the
cv2-backed restatement produced (tools/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from slam_toolkit_b200 import synth
from util import sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(oracle, ex, img, rec):
    assert sha(img) == rec["input"], "synthetic generator drifted"
    k, d = ex.extract(img)
    assert len(k) == rec["n"]
    for l in range(ex.nlevels):
        assert sha(ex.level(l)) == rec["pyramid"][l], f"pyramid level {l}"
        c = ex.candidates(l)
        assert len(c) == rec["ncands"][l] and sha(c) == rec["cands"][l], f"FAST candidates level {l}"
        dd = ex.distributed(l)
        assert len(dd) == rec["ndist"][l] and sha(dd) == rec["dist"][l], f"quadtree level {l}"
        b = ex.blur(l)
        if b is not None:  # the oracle (like the reference) blurs only levels that kept keypoints
            assert sha(b) == rec["blur"][l], f"blur level {l}"
    assert sha(k) == rec["kps"]
    assert sha(d) == rec["desc"]
    return k, d


@pytest.mark.parametrize("seed", range(8))
def test_kitti_seed(oracle, golden, seed):
    ex = oracle.Extractor()
    L, R = synth.stereo_pair(seed)
    g = golden["kitti"][str(seed)]
    kl, dl = _check(oracle, ex, L, g["L"])
    kr, dr = _check(oracle, ex, R, g["R"])
    si, sd = oracle.stereo_match(kl, dl, kr, dr)
    assert sha(si) == g["stereo_idx"] and sha(sd) == g["stereo_dist"]
    assert int((si >= 0).sum()) == g["n_stereo"]


def test_seed0_arrays(oracle):
    z = np.load(os.path.join(ROOT, "tests/golden/golden_seed0.npz"))
    ex = oracle.Extractor()
    L, R = synth.stereo_pair(0)
    kl, dl = ex.extract(L)
    assert np.array_equal(kl, z["kl"]) and np.array_equal(dl, z["dl"])
    kr, dr = ex.extract(R)
    assert np.array_equal(kr, z["kr"]) and np.array_equal(dr, z["dr"])
    # the synthetic pair has uniform disparity 24: accepted matches must show it at level 0
    si = z["stereo_idx"]
    ok = si >= 0
    dx = kl["x"][ok] - kr["x"][si[ok]]
    assert ok.sum() > 800 and np.median(dx) == pytest.approx(24.0, abs=0.5)


@pytest.mark.parametrize("case", range(4))
def test_small_configs(oracle, golden, case):
    rec = golden["small"][str(case)]
    w, h, nf, sf, nl, it, mt = rec["params"]
    ex = oracle.Extractor(nf, sf, nl, it, mt)
    img, _ = synth.stereo_pair(100 + case, w, h)
    k, d = _check(oracle, ex, img, rec)
    z = np.load(os.path.join(ROOT, f"tests/golden/golden_small{case}.npz"))
    assert np.array_equal(k, z["k"]) and np.array_equal(d, z["d"])


def test_tables(oracle):
    t = oracle.Extractor().tables()
    assert t["per_level"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    sizes = [oracle.Extractor().level_size(1241, 376, l) for l in range(8)]
    assert sizes == [(1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151), (416, 126), (346, 105)]


def test_edge_cases(oracle):
    ex = oracle.Extractor()
    k, d = ex.extract(np.zeros((0, 0), np.uint8))
    assert len(k) == 0
    k, d = ex.extract(np.full((376, 1241), 77, np.uint8))  # flat image: no corners anywhere
    assert len(k) == 0 and d.shape == (0, 32)


def test_sequence_with_tracking(oracle, golden):
    """4-frame stereo sequence: the C oracle's extraction equals the cv2 restatement's, and StereoMatch + the tracking
    step (GetDepth + ProjectionMatch) reproduce the committed hashes."""
    g = golden["sequence"]
    L, R = synth.stereo_sequence(5, 4, 4)
    assert [sha(L), sha(R)] == g["inputs"], "synthetic generator drifted"
    cam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0, 0, 0, 0], synth.KITTI_W, synth.KITTI_H)
    ex = oracle.Extractor()
    prev = None
    for f, fr in enumerate(g["frames"]):
        kl, dl = ex.extract(L[f])
        kr, dr = ex.extract(R[f])
        si, _ = oracle.stereo_match(kl, dl, kr, dr)
        assert sha(kl) == fr["kps_l"] and sha(dl) == fr["desc_l"] and sha(si) == fr["stereo_idx"]
        if prev is not None:
            for grid in (False, True):
                ti, td = oracle.track_pair(cam, synth.KITTI_BASELINE, np.eye(4), 50.0, *prev, kl, dl, grid=grid)
                assert sha(ti) == fr["track_idx"] and sha(td) == fr["track_dist"] and int((ti >= 0).sum()) == fr["n_tracked"]
        prev = (kl, dl, kr, si)
