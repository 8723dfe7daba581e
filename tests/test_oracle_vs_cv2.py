"""Pins the dependency-free C oracle to the OpenCV primitives the reference calls
(cv::resize / GaussianBlur / FAST / fastAtan2) and to the cv2-backed Python
restatement of the whole extractor.  CPU only; skipped where cv2 is missing
(the committed fixtures in tests/golden/ then carry the pin)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from hypothesis import given, settings, strategies as st

import oracle_cv2
from slam_toolkit_b200 import synth


@settings(max_examples=25, deadline=None)
@given(st.integers(8, 200), st.integers(8, 120), st.floats(1.05, 2.0), st.integers(0, 2**31))
def test_resize_model(oracle, sw, sh, s, seed):
    src = np.random.default_rng(seed).integers(0, 256, (sh, sw), dtype=np.uint8)
    dw, dh = max(2, int(round(sw / s))), max(2, int(round(sh / s)))
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ref, oracle.resize_linear(src, dw, dh))


@settings(max_examples=25, deadline=None)
@given(st.integers(7, 200), st.integers(7, 120), st.integers(0, 2**31))
def test_gaussian_model(oracle, w, h, seed):
    src = np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
    ref = cv2.GaussianBlur(src.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, oracle.gaussian7(src))


def test_fast_atan2_model(oracle):
    rng = np.random.default_rng(7)
    ys = rng.integers(-1300000, 1300000, 5000)
    xs = rng.integers(-1300000, 1300000, 5000)
    for y, x in list(zip(ys, xs)) + [(0, 0), (0, 5), (5, 0), (-5, 0), (0, -5), (3, 3), (-3, 3)]:
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == oracle.fast_atan2(y, x)


@pytest.mark.parametrize("th", [7, 20, 40])
def test_fast_model(oracle, th):
    rng = np.random.default_rng(th)
    L, _ = synth.stereo_pair(3)
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    total = 0
    for _ in range(200):
        hh, ww = int(rng.integers(5, 44)), int(rng.integers(5, 44))
        y0, x0 = int(rng.integers(0, 376 - hh)), int(rng.integers(0, 1241 - ww))
        cell = np.ascontiguousarray(L[y0:y0 + hh, x0:x0 + ww])
        ref = np.array([(p.pt[0], p.pt[1], p.response) for p in det.detect(cell)], np.float32).reshape(-1, 3)
        got = oracle.fast_nms(cell, th)
        assert np.array_equal(ref, got)
        total += len(ref)
    assert total > 50
    # noise image: many adjacent equal maxima
    noise = rng.integers(0, 256, (60, 60), dtype=np.uint8)
    ref = np.array([(p.pt[0], p.pt[1], p.response) for p in det.detect(noise)], np.float32).reshape(-1, 3)
    assert np.array_equal(ref, oracle.fast_nms(noise, th))


def test_score_map_consistent_with_fast(oracle):
    """best(p) map: corner at t iff best > t, response = best - 1."""
    L, _ = synth.stereo_pair(5)
    cell = np.ascontiguousarray(L[100:160, 300:380])
    s = oracle.fast_score(cell).astype(np.int32)
    det = cv2.FastFeatureDetector_create(7, False, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    pts = {(int(p.pt[0]), int(p.pt[1])) for p in det.detect(cell)}
    mine = {(x, y) for y in range(3, 57) for x in range(3, 77) if s[y, x] > 7}
    assert pts == mine


@pytest.mark.parametrize("seed", [11, 12])
def test_full_extractor_vs_cv2_restatement(oracle, seed):
    img, _ = synth.stereo_pair(seed, 480, 200)
    py = oracle_cv2.ExtractorCv2(600, 1.2, 5, 20, 7)
    stg = {}
    k6, d = py.extract(img, stg)
    ex = oracle.Extractor(600, 1.2, 5, 20, 7)
    k, dd = ex.extract(img)
    for l in range(5):
        assert np.array_equal(ex.level(l), stg["pyramid"][l])
        assert np.array_equal(ex.candidates(l), stg["cands"][l])
        assert np.array_equal(ex.distributed(l), stg["dist"][l])
    got = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(np.float32)], 1)
    assert np.array_equal(got, k6) and np.array_equal(dd, d)


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 600), st.integers(1, 300), st.integers(0, 2**31))
def test_distribute_two_restatements_agree(oracle, n, want, seed):
    """literal std::list emulation (C) vs the list-algebra form (Python), tie rule T1."""
    rng = np.random.default_rng(seed)
    W, H = 400, 110
    pts = set()
    while len(pts) < n:
        pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
    pts = sorted(pts, key=lambda p: (p[1] // 32, p[0] // 32, p[1], p[0]))
    xyr = np.array([(x, y, int(rng.integers(7, 60))) for x, y in pts], np.float32)
    ref = oracle_cv2.ExtractorCv2.distribute([tuple(np.float32(v) for v in r) for r in xyr], 16, 16 + W, 16, 16 + H, want)
    got = oracle.distribute(xyr, 16, 16 + W, 16, 16 + H, want)
    assert np.array_equal(np.array(ref, np.float32).reshape(-1, 3), got)
