"""Pins the C oracle (oracle/orb_oracle.c) to the REFERENCE'S OWN CODE: oracle/_ref = geonuklee/slam-toolkit's
src/orb_extractor.cpp, src/matcher.cpp and src/camera.cpp compiled unmodified (oracle/ref_build/Makefile) against
stand-in third-party headers whose five image primitives are the cv2-pinned models (tests/test_oracle_vs_cv2.py).

With the monotonic heap (list nodes get increasing addresses, so the reference's `sort` by (count, node address),
src/orb_extractor.cpp:684, orders equal counts by creation = the oracle's declared rule T1) every byte must agree:
tables, pyramid, the 19-px ring, FAST candidates, quadtree survivors and their order, angles, blurred planes,
descriptors, StereoMatch, ProjectionMatch.  With glibc's heap the reference itself moves; there the test asserts equality
on every level whose quadtree did NOT stop inside a group of equally-full nodes and reports the rest.
CPU only; skipped when neither the prebuilt library nor /root/reference is present."""
import os

import numpy as np
import pytest

from slam_toolkit_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    import ref_c
    if not ref_c.available():
        pytest.skip("oracle/_ref is not built and the reference sources are absent")
    ref_c.set_heap_mode(ref_c.HEAP_MONOTONIC)
    yield ref_c
    ref_c.set_heap_mode(ref_c.HEAP_MONOTONIC)


def _cam(oracle, d=(0, 0, 0, 0), w=synth.KITTI_W, h=synth.KITTI_H):
    return oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, list(d), w, h)


def _reflect101_ring(img, e=19):
    return np.pad(img, e, mode="reflect")


def _compare_extraction(oracle, ref, o, r, img):
    ko, do = o.extract(img)
    kr, dr = r.extract(img)
    for l in range(o.nlevels):
        lo = o.level(l)
        assert np.array_equal(lo, r.level(l)), f"pyramid level {l}"
        assert np.array_equal(_reflect101_ring(lo), r.level(l, ring=True)), f"reflected ring of level {l}"
        co, cr = o.candidates(l), r.candidates(l)
        assert co.shape == cr.shape and np.array_equal(co, cr), f"FAST candidates level {l} (set and order)"
        w, h = r.level_size(l)
        if len(co):
            quota = int(o.tables()["per_level"][l])
            assert np.array_equal(o.distributed(l), r.distribute(co, 16, w - 16, 16, h - 16, quota, l)), f"quadtree level {l}"
        rb = r.blur(l)
        if rb is not None:
            assert np.array_equal(o.blur(l), rb), f"blurred level {l}"
        else:
            assert len(o.distributed(l)) == 0
    assert len(ko) == len(kr)
    for f in ko.dtype.names:  # angle compared by bits: fastAtan2 output
        assert np.array_equal(ko[f].view(np.uint32), kr[f].view(np.uint32)), f"keypoint field {f}"
    assert np.array_equal(do, dr), "descriptors"
    return ko, do


def test_ctor_tables_equal_the_reference(oracle, ref):
    for nf, sf, nl in [(2000, 1.2, 8), (500, 1.2, 4), (300, 1.5, 3), (1000, 1.2, 8), (50, 1.2, 2), (20, 1.2, 8), (1500, 1.1, 12),
                       (4000, 2.0, 5), (1, 1.2, 8), (777, 1.33, 7)]:
        to, tr = oracle.Extractor(nf, sf, nl, 20, 7).tables(), ref.Extractor(nf, sf, nl, 20, 7).tables()
        for k in to:
            assert np.array_equal(to[k].view(np.uint32), tr[k].view(np.uint32)), (k, nf, sf, nl)


@pytest.mark.parametrize("seed", range(8))
def test_kitti_frames_every_stage(oracle, ref, seed):
    o, r = oracle.Extractor(), ref.Extractor()
    L, R = synth.stereo_pair(seed)
    kl, dl = _compare_extraction(oracle, ref, o, r, L)
    kr, dr = _compare_extraction(oracle, ref, o, r, R)
    si, _ = oracle.stereo_match(kl, dl, kr, dr)
    assert np.array_equal(si, ref.stereo_match(kl, dl, kr, dr, _cam(oracle)))
    assert (si >= 0).sum() > 1000


@pytest.mark.parametrize("case", [(320, 240, 500, 1.2, 4, 20, 7), (161, 131, 300, 1.5, 3, 20, 7), (640, 200, 1000, 1.2, 8, 30, 10),
                                  (97, 95, 50, 1.2, 2, 20, 7), (1241, 376, 20, 1.2, 8, 20, 7), (800, 600, 3000, 1.2, 8, 20, 7),
                                  (333, 500, 700, 1.3, 6, 25, 5), (1920, 1080, 2000, 1.2, 8, 20, 7)])
def test_other_geometries(oracle, ref, case):
    w, h, nf, sf, nl, it, mt = case
    img, _ = synth.stereo_pair(100 + w % 7, w, h)
    _compare_extraction(oracle, ref, oracle.Extractor(nf, sf, nl, it, mt), ref.Extractor(nf, sf, nl, it, mt), img)


def test_noise_flat_and_retry_images(oracle, ref):
    """noise (every level far over quota, very long careful phase), flat (nothing), low contrast (20 -> 7 retry)"""
    rng = np.random.default_rng(3)
    o, r = oracle.Extractor(), ref.Extractor()
    noise = rng.integers(0, 256, (376, 1241), dtype=np.uint8)
    k, _ = _compare_extraction(oracle, ref, o, r, noise)
    assert len(k) >= 2000
    k, _ = _compare_extraction(oracle, ref, o, r, np.full((376, 1241), 77, np.uint8))
    assert len(k) == 0
    soft = (synth.stereo_pair(11)[0].astype(np.float32) * 0.12 + 100).astype(np.uint8)   # contrast below iniThFAST
    k, _ = _compare_extraction(oracle, ref, o, r, soft)
    assert len(k) > 0
    mixed = synth.stereo_pair(12)[0].copy()
    mixed[:, 600:] = (mixed[:, 600:].astype(np.float32) * 0.1 + 90).astype(np.uint8)
    _compare_extraction(oracle, ref, o, r, mixed)


def test_distribute_random_candidate_sets(oracle, ref):
    """DistributeOctTree + DivideNode (src/orb_extractor.cpp:481-763) on synthetic candidate lists with many equal counts"""
    r = ref.Extractor()
    rng = np.random.default_rng(5)
    for t in range(60):
        W, H = int(rng.integers(40, 1300)), int(rng.integers(40, 500))
        if round(W / H) < 1:
            continue
        n = int(rng.integers(1, 4000))
        xs = rng.integers(0, W - 3, n).astype(np.float32)
        ys = rng.integers(0, H - 3, n).astype(np.float32)
        if t % 3 == 0:   # duplicates and clusters
            xs[: n // 2] = xs[0]
            ys[: n // 3] = ys[0]
        resp = rng.integers(8, 120, n).astype(np.float32)
        xyr = np.stack([xs, ys, resp], 1)
        want = int(rng.integers(1, 2500))
        a = oracle.distribute(xyr, 16, 16 + W, 16, 16 + H, want)
        b = r.distribute(xyr, 16, 16 + W, 16, 16 + H, want)
        assert np.array_equal(a, b), (t, W, H, n, want)


def test_glibc_heap_moves_only_levels_cut_inside_a_tie_group(oracle, ref):
    """The reference as it runs on this machine (glibc malloc): equal to the oracle as a SET on every level whose careful
    phase did not stop between two equally-full nodes; the others are where the heap-address order decides (T1)."""
    o, r = oracle.Extractor(), ref.Extractor()
    total = moved = cut_levels = levels = 0
    for seed in range(8):
        img, _ = synth.stereo_pair(seed)
        ref.set_heap_mode(ref.HEAP_MALLOC)
        kr, dr = r.extract(img)
        ref.set_heap_mode(ref.HEAP_MONOTONIC)
        ko, do = o.extract(img)
        for l in range(8):
            so = {kp.tobytes() + d.tobytes() for kp, d in zip(ko[ko["octave"] == l], do[ko["octave"] == l])}
            sr = {kp.tobytes() + d.tobytes() for kp, d in zip(kr[kr["octave"] == l], dr[kr["octave"] == l])}
            levels += 1
            if o.tie_cut(l):
                cut_levels += 1
                moved += len(so - sr)
            else:
                assert so == sr, f"seed {seed} level {l}: no tie at the cut, yet the keypoint sets differ"
        total += len(ko)
    print(f"glibc heap: {moved} of {total} keypoints differ, all on the {cut_levels} of {levels} levels cut inside a tie group")
    assert moved < 0.05 * total


def test_descriptor_distance_and_atan2(oracle, ref):
    rng = np.random.default_rng(9)
    d = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    for i in range(0, 300, 2):
        assert oracle.hamming256(d[i], d[i + 1]) == ref.hamming256(d[i], d[i + 1]) == int(np.unpackbits(d[i] ^ d[i + 1]).sum())


def test_se3_and_camera_primitives(oracle, ref):
    """predicted_Tcw * Xw (g2o::SE3Quat, quaternion product), Camera::Project / IsInImage / NormalizedUndistort by bits"""
    rng = np.random.default_rng(0)
    for t in range(100):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if t % 3 == 0:
            q = np.array([0, 0, 0, 1.0]) + rng.normal(size=4) * 1e-2
        qt = np.concatenate([q, rng.normal(size=3)])
        x = rng.normal(size=(300, 3)) * rng.uniform(0.1, 100)
        yr, qh = ref.se3_apply(qt, x)
        assert abs(np.linalg.norm(qh) - 1) < 1e-15 and qh[3] >= 0
        yo = oracle.se3_apply(np.concatenate([qh, qt[4:]]), x)
        assert np.array_equal(yr.view(np.uint64), yo.view(np.uint64))
    cam = _cam(oracle, (-0.1, 0.02, 0.001, -0.0005))
    kps = np.zeros(500, oracle.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, 1241, 500), rng.uniform(0, 376, 500)
    assert np.array_equal(oracle.normalized_undistort(cam, kps).view(np.uint64), ref.normalized_undistort(cam, kps).view(np.uint64))


def _projection_inputs(seed, kl, dl, n=3000):
    rng = np.random.default_rng(seed)
    z = rng.uniform(2, 80, n)
    u, v = rng.uniform(-50, 1291, n), rng.uniform(-30, 406, n)
    xw = np.stack([(u - synth.KITTI_CX) / synth.KITTI_FX * z, (v - synth.KITTI_CY) / synth.KITTI_FY * z, z], 1)
    xw[::17, 2] *= -1
    md = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    pick = rng.integers(0, len(kl), n)
    cp = rng.random(n) < 0.5
    md[cp] = dl[pick[cp]]
    for i in np.nonzero(cp)[0]:
        for b in rng.integers(0, 256, 3):
            md[i, b // 8] ^= np.uint8(1 << (b % 8))
    # two queries with the same descriptor landing on one keypoint: equal distance, the later one must win (:197-204)
    md[n - 1], xw[n - 1] = md[np.nonzero(cp)[0][0]], xw[np.nonzero(cp)[0][0]]
    skip = (rng.random(n) < 0.05).astype(np.uint8)
    return xw, md, skip


@pytest.mark.parametrize("seed", range(3))
def test_projection_match_equals_the_reference(oracle, ref, seed):
    r = ref.Extractor()
    kl, dl = r.extract(synth.stereo_pair(seed)[0])
    xw, md, skip = _projection_inputs(seed, kl, dl)
    for d4 in ((0, 0, 0, 0), (-0.1, 0.02, 0.001, -0.0005)):
        cam = _cam(oracle, d4)
        for qt_in in ([0, 0, 0, 1, 0, 0, 0], [0.01, -0.02, 0.005, 0.9997, 0.1, -0.05, 0.2], [0.3, -0.1, 0.2, -0.9, 1.0, 0.5, 4.0]):
            _, q = ref.se3_apply(np.array(qt_in, float), xw[:1])
            qt = np.concatenate([q, np.array(qt_in[4:], float)])
            for radius in (50.0, 10.0, 100.0):
                want = ref.projection_match(xw, md, skip, qt, cam, kl, dl, radius)
                for grid in (False, True):
                    got, dist = oracle.projection_match(xw, md, skip, qt, cam, kl, dl, radius, grid=grid)
                    assert np.array_equal(got, want), (d4, qt_in, radius, grid)
                    ok = got >= 0
                    assert all(dist[j] == oracle.hamming256(md[got[j]], dl[j]) for j in np.nonzero(ok)[0])


def test_projection_golden_fixture(oracle):
    """the committed ProjectionMatch result of the reference (tools/gen_golden.py) -- runs without _ref (GPU box)"""
    z = np.load(os.path.join(ROOT, "tests/golden/golden_proj.npz"))
    g = np.load(os.path.join(ROOT, "tests/golden/golden_seed0.npz"))
    cam = _cam(oracle, z["dist4"])
    assert np.array_equal(oracle.se3_apply(z["qt"], z["xw"]).view(np.uint64), z["xc"].view(np.uint64))
    for radius in (50, 10):
        got, _ = oracle.projection_match(z["xw"], z["mp_desc"], z["skip"], z["qt"], cam, g["kl"], g["dl"], float(radius))
        assert np.array_equal(got, z[f"to_query_r{radius}"])
        assert (got >= 0).sum() > 10


def test_stereo_match_edge_cases_equal_the_reference(oracle, ref):
    """hand-built rows: single candidate (accepted, dist1 = 999999999), tie for best (rejected), dy = +-3 exactly,
    dx = 0 / 100 / just outside, negative dx, bucket borders (y multiples of 10), an empty right set"""
    rng = np.random.default_rng(21)
    cam = _cam(oracle)
    for trial in range(30):
        nl, nr = int(rng.integers(1, 200)), int(rng.integers(0, 200))
        kl, kr = np.zeros(nl, oracle.KP_DTYPE), np.zeros(nr, oracle.KP_DTYPE)
        kl["x"], kl["y"] = rng.integers(0, 1241, nl), rng.integers(0, 40, nl)   # few rows: many candidates per keypoint
        kr["x"] = rng.integers(0, 1241, nr)
        kr["y"] = rng.integers(0, 40, nr) + rng.choice([0.0, 0.5, -0.25], nr)
        dl = rng.integers(0, 256, (nl, 32), dtype=np.uint8)
        dr = rng.integers(0, 256, (nr, 32), dtype=np.uint8)
        if nr > 4:
            dr[1] = dr[0]                               # exact tie -> rejected wherever both are candidates
            kr[1] = kr[0]
            kl[0]["x"], kl[0]["y"] = kr[0]["x"] + 100, kr[0]["y"] + 3     # on both limits
            dl[0] = dr[2]
            kr[2]["x"], kr[2]["y"] = kl[0]["x"], kl[0]["y"] - 3           # dx = 0, dy = 3
        got, _ = oracle.stereo_match(kl, dl, kr, dr)
        assert np.array_equal(got, ref.stereo_match(kl, dl, kr, dr, cam)), trial
