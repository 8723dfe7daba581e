"""Parity tests proper: the CUDA path through the C ABI (libsfe.so) against the CPU oracle and the
committed golden fixtures.  Integer / byte / index work is compared bit-exactly; the float fields
of a keypoint (x, y, size, angle, response) are compared bit-exactly too (tolerance stated by the
north star: angle 1e-4 rad, pyramid +-1 LSB -- we hold 0)."""
import os

import numpy as np
import pytest

from slam_toolkit_b200 import api, synth
from util import sha

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu():
    if api.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return 0


@pytest.fixture(scope="module")
def kitti_ex(gpu):
    return api.ORBextractor(max_images=8)


def _stage_report(ex, ref, image_idx, nlevels, cap=1 << 15):
    bad = []
    for l in range(nlevels):
        if not np.array_equal(ex.debug_level(image_idx, l), ref.level(l)):
            bad.append(f"pyramid L{l}: {(ex.debug_level(image_idx, l) != ref.level(l)).sum()} px differ")
        c, rc = ex.debug_points(image_idx, l, cap=cap), ref.candidates(l, cap=cap)
        if not np.array_equal(c, rc):
            bad.append(f"FAST candidates L{l}: {len(c)} vs {len(rc)}")
        d, rd = ex.debug_points(image_idx, l, distributed=True, cap=cap), ref.distributed(l, cap=cap)
        if not np.array_equal(d, rd):
            bad.append(f"quadtree L{l}: {len(d)} vs {len(rd)}")
        rb = ref.blur(l)
        if rb is not None and not np.array_equal(ex.debug_level(image_idx, l, blur=True), rb):
            bad.append(f"blur L{l}: {(ex.debug_level(image_idx, l, blur=True) != rb).sum()} px differ")
    return bad


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_extract_kitti_vs_oracle_and_golden(kitti_ex, oracle, golden, seed):
    L, _ = synth.stereo_pair(seed)
    k, d = kitti_ex.extract(L)
    ref = oracle.Extractor()
    rk, rd = ref.extract(L)
    stages = _stage_report(kitti_ex, ref, 0, 8)
    assert not stages, stages
    assert len(k) == len(rk)
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(k[f], rk[f]), f"keypoint field {f}: {(k[f] != rk[f]).sum()} differ"
    assert np.array_equal(d, rd), f"{(d != rd).any(axis=1).sum()} descriptors differ"
    g = golden["kitti"][str(seed)]["L"]
    assert sha(k) == g["kps"] and sha(d) == g["desc"]


@pytest.mark.parametrize("case", range(4))
def test_extract_small_configs(gpu, oracle, golden, case):
    rec = golden["small"][str(case)]
    w, h, nf, sf, nl, it, mt = rec["params"]
    img, _ = synth.stereo_pair(100 + case, w, h)
    ex = api.ORBextractor(nf, sf, nl, it, mt)
    k, d = ex.extract(img)
    ref = oracle.Extractor(nf, sf, nl, it, mt)
    rk, rd = ref.extract(img)
    stages = _stage_report(ex, ref, 0, nl)
    assert not stages, stages
    assert np.array_equal(k, rk) and np.array_equal(d, rd)
    assert sha(k) == rec["kps"] and sha(d) == rec["desc"]


def test_edge_images(kitti_ex):
    k, d = kitti_ex.extract(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = kitti_ex.extract(np.full((376, 1241), 90, np.uint8))  # flat: no corner anywhere
    assert len(k) == 0
    # a FAST cell with < 7 rows, a level whose window is < 30 px (T5), strided input
    img, _ = synth.stereo_pair(42, 200, 95)
    ex = api.ORBextractor(200, 1.3, 5, 20, 7)
    k, d = ex.extract(img)
    import oracle_c
    rk, rd = oracle_c.Extractor(200, 1.3, 5, 20, 7).extract(img)
    assert np.array_equal(k, rk) and np.array_equal(d, rd)


def test_fallback_threshold_cells(gpu, oracle):
    """low-contrast image: most cells find nothing at 20 and retry at 7 (src/orb_extractor.cpp:811-816)."""
    img, _ = synth.stereo_pair(7, 400, 200)
    img = (96 + (img.astype(np.int32) - 96) // 3).astype(np.uint8)
    ex = api.ORBextractor(500, 1.2, 4, 20, 7)
    k, d = ex.extract(img)
    ref = oracle.Extractor(500, 1.2, 4, 20, 7)
    rk, rd = ref.extract(img)
    assert len(rk) > 50 and (rk["response"] < 20).any()
    assert not _stage_report(ex, ref, 0, 4)
    assert np.array_equal(k, rk) and np.array_equal(d, rd)


def test_fast_segments_mixed_retry_noise_and_high_thresholds(gpu, oracle):
    """FAST by segments: a segment whose cells partly retry at minThFAST (the neighbours must not see each other's scores),
    a noise image where most pixels survive the compass pre-test (list capacity = every tested pixel), cells 32 px wide
    (3 cells per segment) and thresholds above 127 (the VABSDIFF4 pre-test clamps its threshold there)."""
    img, _ = synth.stereo_pair(9, 500, 230)
    mixed = img.copy()
    mixed[:, 130:330] = (96 + (mixed[:, 130:330].astype(np.int32) - 96) // 4).astype(np.uint8)   # low-contrast stripe: retry cells
    rng = np.random.default_rng(5)
    noise = np.clip(rng.normal(128, 22, (230, 500)), 0, 255).astype(np.uint8)
    contrast = ((img.astype(np.int32) > 110) * 255).astype(np.uint8)
    for name, im, prm in (("mixed", mixed, (600, 1.2, 4, 20, 7)), ("noise", noise, (400, 1.2, 3, 20, 7)),
                          ("wide cells", img[:, :470], (500, 1.25, 4, 20, 7)), ("t>127", contrast, (300, 1.2, 3, 150, 130)),
                          ("ini<min", img, (300, 1.2, 3, 10, 25))):
        ex = api.ORBextractor(*prm)
        ref = oracle.Extractor(*prm)
        rk, rd = ref.extract(np.ascontiguousarray(im))
        k, d = ex.extract(np.ascontiguousarray(im))
        bad = _stage_report(ex, ref, 0, prm[2])
        assert not bad, (name, bad)
        assert np.array_equal(k, rk) and np.array_equal(d, rd), name
        assert len(rk) > 20, (name, len(rk))
    rk, _ = oracle.Extractor(600, 1.2, 4, 20, 7).extract(mixed)
    assert (rk["response"] < 20).any() and (rk["response"] >= 20).any()


@pytest.mark.parametrize("w,h,nf,sf,nl", [(1920, 1080, 3000, 1.2, 8), (752, 480, 1200, 1.5, 5), (640, 480, 1000, 2.0, 4),
                                           (1024, 300, 1500, 1.1, 10), (333, 217, 400, 1.7, 3), (2048, 256, 2000, 1.25, 6)])
def test_other_geometries_and_scale_factors(gpu, oracle, w, h, nf, sf, nl):
    """Other camera geometries and pyramid scale factors through the whole extractor: cell widths / segment packing,
    the 4-outputs-per-thread pyramid pass (source pairs must fit 8 bytes: scale <= 2.33) and its fallback, TMA boxes."""
    img, right = synth.stereo_pair(31, w, h)
    ex = api.ORBextractor(nf, sf, nl, 20, 7, max_images=2)
    ref = oracle.Extractor(nf, sf, nl, 20, 7)
    out = ex.stereo_frames(img[None], right[None])
    rk, rd = ref.extract(img)
    bad = _stage_report(ex, ref, 0, nl)
    assert not bad, bad
    n = out["n_l"][0]
    assert n == len(rk) and np.array_equal(out["kps_l"][0, :n], rk) and np.array_equal(out["desc_l"][0, :n], rd)
    rkr, rdr = ref.extract(right)
    si, sd = oracle.stereo_match(rk, rd, rkr, rdr)
    assert np.array_equal(out["stereo_idx"][0, :n], si) and np.array_equal(out["stereo_dist"][0, :n], sd)
    assert len(rk) > nf // 3


def test_sequence_edge_cases_and_replanning(gpu, oracle):
    """A flat frame in the middle of a sequence (zero keypoints: nothing to track from or to), a one-frame sequence, few and
    many requested features, and the same handle re-planned for another image size between calls."""
    w, h = 480, 200
    L, R = synth.stereo_sequence(3, 4, 4, w, h)
    L[2] = 77
    R[2] = 77
    cam = api.Camera.make(300.0, 300.0, w / 2.0, h / 2.0, (0.01, -0.001, 0, 0), w, h)
    tp = api.TrackParams.make(cam, 0.4, None, 25.0)
    ex = api.ORBextractor(300, 1.2, 4, 20, 7, max_images=8)
    out = ex.stereo_sequence(L, R, tp)
    assert out["n_l"][2] == 0 and out["n_r"][2] == 0 and out["n_l"][1] > 100
    assert (out["track_idx"][2] == -1).all() and (out["track_idx"][3] == -1).all()   # nothing in frame 2, nothing from it
    tracked = _check_tracking(oracle, out, 4, tp, np.eye(4))
    assert tracked > 50
    one = ex.stereo_sequence(L[:1], R[:1], tp)
    assert (one["track_idx"] == -1).all() and one["n_l"][0] == out["n_l"][0]
    assert np.array_equal(one["kps_l"][0], out["kps_l"][0])
    # re-plan the same handle for another size, then back
    img2, _ = synth.stereo_pair(8, 301, 177)
    k2, d2 = ex.extract(img2)
    rk2, rd2 = oracle.Extractor(300, 1.2, 4, 20, 7).extract(img2)
    assert np.array_equal(k2, rk2) and np.array_equal(d2, rd2)
    again = ex.stereo_sequence(L, R, tp)
    _stereo_equal(again, out, 4)   # (rows past a frame's keypoint count are not defined)
    assert np.array_equal(again["track_idx"], out["track_idx"]) and np.array_equal(again["track_dist"], out["track_dist"])
    for nf in (20, 6000):
        img, _ = synth.stereo_pair(12, 900, 400)
        e2 = api.ORBextractor(nf, 1.2, 5, 20, 7)
        k, d = e2.extract(img)
        rk, rd = oracle.Extractor(nf, 1.2, 5, 20, 7).extract(img)
        assert np.array_equal(k, rk) and np.array_equal(d, rd), nf
        assert len(rk) >= nf if nf == 20 else len(rk) > 3000


def test_4k_frames_dense_and_noise_images(gpu, oracle):
    """3840 x 2160: coordinates beyond the stereo matcher's bucket grid (clamped buckets), a tracking bucket grid that
    needs the opt-in shared-memory size, wide levels, 44 k corners on one level.  A noise frame with 840 k FAST corners on
    level 0 -- far beyond any pre-sized candidate buffer and beyond the quadtree's 16-bit instance -- is NOT refused (the
    reference's candidate list is unbounded, src/orb_extractor.cpp:778-779): the call re-runs itself with buffers sized
    from its own count and returns the reference's keypoints."""
    w, h = 3840, 2160
    Ls, Rs = synth.stereo_sequence(2, 2, 4, 1280, 720)

    def up(a):   # 3x upsampling + 3x3 box filter: 4K frames with a moderate corner density (~10 k on level 0)
        b = np.pad(np.repeat(np.repeat(a.astype(np.float32), 3, axis=0), 3, axis=1), 1, mode="edge")
        s = sum(b[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)) / 9.0
        return np.clip(np.rint(s), 0, 255).astype(np.uint8)
    L, R = np.stack([up(x) for x in Ls]), np.stack([up(x) for x in Rs])
    cam = api.Camera.make(2100.0, 2100.0, w / 2.0, h / 2.0, (0, 0, 0, 0), w, h)
    tp = api.TrackParams.make(cam, 0.5, None, 60.0)
    ex = api.ORBextractor(5000, 1.2, 8, 20, 7, max_images=4)
    ref = oracle.Extractor(5000, 1.2, 8, 20, 7)
    rk, rd = ref.extract(L[0])
    assert 3072 < max(len(ref.candidates(l)) for l in range(8))
    out = ex.stereo_sequence(L, R, tp)
    n = out["n_l"][0]
    assert n == len(rk) and np.array_equal(out["kps_l"][0, :n], rk) and np.array_equal(out["desc_l"][0, :n], rd)
    rkr, rdr = ref.extract(R[0])
    si, sd = oracle.stereo_match(rk, rd, rkr, rdr)
    assert np.array_equal(out["stereo_idx"][0, :n], si) and np.array_equal(out["stereo_dist"][0, :n], sd)
    assert (si >= 0).sum() > 500
    assert _check_tracking(oracle, out, 2, tp, np.eye(4)) > 200
    dense, _ = synth.stereo_pair(2, w, h)     # ~44 k corners on level 0: every level runs the quadtree on global scratch
    dk, dd = ref.extract(dense)
    assert max(len(ref.candidates(l)) for l in range(8)) > 40000
    k, d = ex.extract(dense)
    assert np.array_equal(k, dk) and np.array_equal(d, dd)
    noise = np.random.default_rng(1).integers(0, 256, (h, w), dtype=np.uint8)   # 840 k / 533 k / 332 k ... corners per level
    nk, nd = ref.extract(noise)
    assert len(ref.candidates(0, cap=1 << 21)) > 800000
    k, d = ex.extract(noise)                  # first pass overflows, the call re-runs with the 32-bit quadtree
    assert len(k) == len(nk) >= 5000 and np.array_equal(k, nk) and np.array_equal(d, nd)
    assert _stage_report(ex, ref, 0, 8, cap=1 << 21) == []
    k, d = ex.extract(L[0])                   # and ordinary frames still match afterwards
    assert np.array_equal(k, rk) and np.array_equal(d, rd)


def test_candidate_overflow_reruns_instead_of_failing(gpu, oracle, monkeypatch):
    """Tiny initial candidate buffers (SFE_CAND_CAP=64): every entry point -- single image, pipelined host batch, resident
    batch -- re-runs with room for what the first pass counted and returns the oracle's bytes; an asynchronous resident
    queue reports SFE_ERR_CAPACITY at wait() (its inputs belong to the caller) and succeeds when the batch is re-submitted."""
    monkeypatch.setenv("SFE_CAND_CAP", "64")
    ex = api.ORBextractor(max_images=8)
    ref = oracle.Extractor()
    imgs = np.stack([synth.stereo_pair(s)[i] for s in (20, 21) for i in (0, 1)])
    want = [ref.extract(im) for im in imgs]
    k, d = ex.extract(imgs[0])
    assert np.array_equal(k, want[0][0]) and np.array_equal(d, want[0][1])
    monkeypatch.setenv("SFE_PIPELINE_CHUNKS", "2")
    ex2 = api.ORBextractor(max_images=8)
    kps, desc, n = ex2.extract_batch(imgs)
    for i, (wk, wd) in enumerate(want):
        assert n[i] == len(wk) and np.array_equal(kps[i, :n[i]], wk) and np.array_equal(desc[i, :n[i]], wd)
    out = ex2.stereo_frames(imgs[0::2], imgs[1::2])
    for f in range(2):
        assert np.array_equal(out["kps_r"][f, :out["n_r"][f]], want[2 * f + 1][0])
    # resident, synchronous and asynchronous
    ex3 = api.ORBextractor(max_images=8)
    H, W = imgs.shape[1:]
    d_img = api.DeviceBuffer(imgs.nbytes)
    d_img.upload(imgs)
    cap = ex3.cap
    d_k, d_d, d_n = api.DeviceBuffer(4 * cap * 28), api.DeviceBuffer(4 * cap * 32), api.DeviceBuffer(16)
    ex3.extract_batch_dev(d_img.ptr, 4, W, H, d_k.ptr, d_d.ptr, d_n.ptr)
    n = d_n.download((4,), np.int32)
    kk = d_k.download((4, cap), api.KP_DTYPE)
    for i, (wk, _) in enumerate(want):
        assert n[i] == len(wk) and np.array_equal(kk[i, :n[i]], wk)
    ex4 = api.ORBextractor(max_images=8)
    ex4.set_async(True)
    ex4.extract_batch_dev(d_img.ptr, 4, W, H, d_k.ptr, d_d.ptr, d_n.ptr)
    with pytest.raises(api.SfeError) as err:
        ex4.wait()
    assert err.value.status == api.SFE_ERR_CAPACITY and "submit the batch again" in str(err.value)
    ex4.extract_batch_dev(d_img.ptr, 4, W, H, d_k.ptr, d_d.ptr, d_n.ptr)
    ex4.wait()
    n = d_n.download((4,), np.int32)
    kk = d_k.download((4, cap), api.KP_DTYPE)
    for i, (wk, _) in enumerate(want):
        assert n[i] == len(wk) and np.array_equal(kk[i, :n[i]], wk)


def test_few_features_on_a_wide_image(gpu, oracle):
    """nfeatures = 20 on a KITTI frame: a level keeps all 4 * nIni = 16 first-split nodes although its quota is 2-4
    (src/orb_extractor.cpp:606-669), so the frame returns far more than nfeatures + a few per level"""
    img = synth.stereo_pair(3)[0]
    ex = api.ORBextractor(20, 1.2, 8, 20, 7)
    rk, rd = oracle.Extractor(20, 1.2, 8, 20, 7).extract(img)
    assert len(rk) > 20 + 4 * 8
    assert ex.cap_for(1241, 376) >= len(rk)
    k, d = ex.extract(img)
    assert np.array_equal(k, rk) and np.array_equal(d, rd)


def test_ctor_tables_equal_the_oracle(gpu, oracle):
    """GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares (the last is the one
    the reference reads, src/pipeline.cpp:89) and the per-level quotas, by float bits"""
    for nf, sf, nl in [(2000, 1.2, 8), (500, 1.2, 4), (300, 1.5, 3), (1500, 1.1, 12), (4000, 2.0, 5), (777, 1.33, 7)]:
        ex = api.ORBextractor(nf, sf, nl, 20, 7)
        t = oracle.Extractor(nf, sf, nl, 20, 7).tables()
        for got, key in zip((ex.GetScaleFactors(), ex.GetInverseScaleFactors(), ex.GetScaleSigmaSquares(),
                             ex.GetInverseScaleSigmaSquares()), ("scale", "inv_scale", "sigma2", "inv_sigma2")):
            assert np.array_equal(np.asarray(got, np.float32).view(np.uint32), t[key].view(np.uint32)), (key, nf, sf, nl)
        assert np.array_equal(np.asarray(ex.features_per_level(), np.int32), t["per_level"])


def test_batch_equals_single(kitti_ex, oracle):
    imgs = np.stack([synth.stereo_pair(s)[i] for s in (4, 5) for i in (0, 1)])
    kps, desc, n = kitti_ex.extract_batch(imgs)
    ref = oracle.Extractor()
    for i in range(4):
        rk, rd = ref.extract(imgs[i])
        assert n[i] == len(rk)
        assert np.array_equal(kps[i, :n[i]], rk) and np.array_equal(desc[i, :n[i]], rd), f"image {i}"


def test_stereo_frames_vs_oracle_and_golden(kitti_ex, oracle, golden):
    seeds = [0, 6, 7]
    L = np.stack([synth.stereo_pair(s)[0] for s in seeds])
    R = np.stack([synth.stereo_pair(s)[1] for s in seeds])
    out = kitti_ex.stereo_frames(L, R)
    ref = oracle.Extractor()
    for f, s in enumerate(seeds):
        kl, dl = ref.extract(L[f])
        kr, dr = ref.extract(R[f])
        si, sd = oracle.stereo_match(kl, dl, kr, dr)
        nl, nr = out["n_l"][f], out["n_r"][f]
        assert (nl, nr) == (len(kl), len(kr))
        assert np.array_equal(out["kps_l"][f, :nl], kl) and np.array_equal(out["desc_l"][f, :nl], dl)
        assert np.array_equal(out["kps_r"][f, :nr], kr) and np.array_equal(out["desc_r"][f, :nr], dr)
        assert np.array_equal(out["stereo_idx"][f, :nl], si), f"{(out['stereo_idx'][f, :nl] != si).sum()} stereo indices differ"
        assert np.array_equal(out["stereo_dist"][f, :nl], sd)
        g = golden["kitti"][str(s)]
        assert sha(si) == g["stereo_idx"] and int((si >= 0).sum()) == g["n_stereo"]
        # domain property: uniform disparity 24 px on level-0 matches
        ok = (si >= 0) & (kl["octave"] == 0)
        assert np.median(kl["x"][ok] - kr["x"][si[ok]]) == 24.0


def _kitti_track_params(radius=50.0, Tcw=None):
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, (0, 0, 0, 0), synth.KITTI_W, synth.KITTI_H)
    return api.TrackParams.make(cam, synth.KITTI_BASELINE, Tcw, radius)


def _check_tracking(oracle, out, frames, tp, rt):
    ocam = oracle.make_camera(tp.cam.fx, tp.cam.fy, tp.cam.cx, tp.cam.cy, list(tp.cam.d), tp.cam.width, tp.cam.height)
    assert (out["track_idx"][0] == -1).all() and (out["track_dist"][0] == -1).all()
    tracked = 0
    for f in range(1, frames):
        nl, npv, nrp = out["n_l"][f], out["n_l"][f - 1], out["n_r"][f - 1]
        ti, td = oracle.track_pair(ocam, tp.baseline, rt, tp.radius, out["kps_l"][f - 1, :npv], out["desc_l"][f - 1, :npv],
                                   out["kps_r"][f - 1, :nrp], out["stereo_idx"][f - 1, :npv], out["kps_l"][f, :nl],
                                   out["desc_l"][f, :nl])
        assert np.array_equal(out["track_idx"][f, :nl], ti), f"frame {f}: {(out['track_idx'][f, :nl] != ti).sum()} track indices differ"
        assert np.array_equal(out["track_dist"][f, :nl], td)
        assert (out["track_idx"][f, nl:] == -1).all()
        tracked += int((ti >= 0).sum())
    return tracked


def test_stereo_sequence_vs_golden(kitti_ex, golden):
    """The committed hashes of the 4-frame sequence (tools/gen_golden.py: cv2 restatement + matcher oracle)."""
    g = golden["sequence"]
    L, R = synth.stereo_sequence(5, 4, 4)
    assert [sha(L), sha(R)] == g["inputs"]
    out = kitti_ex.stereo_sequence(L, R, _kitti_track_params())
    for f, fr in enumerate(g["frames"]):
        n = out["n_l"][f]
        assert sha(out["kps_l"][f, :n]) == fr["kps_l"] and sha(out["desc_l"][f, :n]) == fr["desc_l"]
        assert sha(out["stereo_idx"][f, :n]) == fr["stereo_idx"]
        if f:
            assert sha(out["track_idx"][f, :n]) == fr["track_idx"] and sha(out["track_dist"][f, :n]) == fr["track_dist"]
            assert int((out["track_idx"][f, :n] >= 0).sum()) == fr["n_tracked"]


def test_stereo_sequence_tracking_vs_oracle(kitti_ex, oracle, monkeypatch):
    """sfe_stereo_sequence: the extraction / stereo outputs equal sfe_stereo_frames', and every frame's track_idx equals
    GetDepth + ProjectionMatch of the oracle on the previous frame (host path, pipelined host path, resident path)."""
    F = 4
    L, R = synth.stereo_sequence(5, F, step=4)
    tp = _kitti_track_params()
    plain = kitti_ex.stereo_frames(L, R)
    out = kitti_ex.stereo_sequence(L, R, tp)
    _stereo_equal(out, plain, F)
    tracked = _check_tracking(oracle, out, F, tp, np.eye(4))
    # the camera slid 4 px between frames: most stereo points of a frame are found again in the next one
    assert tracked > 0.5 * int((out["stereo_idx"][:F - 1] >= 0).sum())
    ok = out["track_idx"][1] >= 0
    dxs = out["kps_l"][0]["x"][out["track_idx"][1][ok]] - out["kps_l"][1]["x"][ok]
    assert np.median(dxs[out["kps_l"][1]["octave"][ok] == 0]) == 4.0
    # sub-batches of one frame each: consecutive frames live in different chunks
    monkeypatch.setenv("SFE_PIPELINE_CHUNKS", "4")
    ex2 = api.ORBextractor(max_images=8)
    out2 = ex2.stereo_sequence(L, R, tp)
    _stereo_equal(out2, out, F)
    assert np.array_equal(out2["track_idx"], out["track_idx"]) and np.array_equal(out2["track_dist"], out["track_dist"])
    monkeypatch.delenv("SFE_PIPELINE_CHUNKS")
    # resident entry point, a motion prior that is not the identity, a small radius
    Tcw = np.eye(4)
    Tcw[0, 3], Tcw[2, 3] = -0.05, 0.02
    tp2 = _kitti_track_params(radius=12.0, Tcw=Tcw)
    cap, h, w = kitti_ex.cap, *L.shape[1:]
    dl, dr = api.DeviceBuffer(L.nbytes).upload(L), api.DeviceBuffer(R.nbytes).upload(R)
    spec = {"kps_l": (28, api.KP_DTYPE, (F, cap)), "desc_l": (32, np.uint8, (F, cap, 32)), "n_l": (0, np.int32, (F,)),
            "kps_r": (28, api.KP_DTYPE, (F, cap)), "desc_r": (32, np.uint8, (F, cap, 32)), "n_r": (0, np.int32, (F,)),
            "stereo_idx": (4, np.int32, (F, cap)), "stereo_dist": (4, np.int32, (F, cap)), "track_idx": (4, np.int32, (F, cap)),
            "track_dist": (4, np.int32, (F, cap))}
    bufs = {k: api.DeviceBuffer(max(b * cap * F, 4 * F)) for k, (b, _, _) in spec.items()}
    kitti_ex.stereo_sequence_dev(dl.ptr, dr.ptr, F, w, h, {k: b.ptr for k, b in bufs.items()}, tp2)
    dev = {k: bufs[k].download(shape, dt) for k, (_, dt, shape) in spec.items()}
    _stereo_equal(dev, plain, F)
    _check_tracking(oracle, dev, F, tp2, Tcw)


def test_resident_entry_points_match_host_ones(kitti_ex, oracle):
    L, R = synth.stereo_pair(3)
    cap, h, w = kitti_ex.cap, *L.shape
    host = kitti_ex.stereo_frames(L[None], R[None])
    dl, dr = api.DeviceBuffer(L.nbytes).upload(L), api.DeviceBuffer(R.nbytes).upload(R)
    spec = {"kps_l": 28 * cap, "desc_l": 32 * cap, "n_l": 4, "kps_r": 28 * cap, "desc_r": 32 * cap, "n_r": 4,
            "stereo_idx": 4 * cap, "stereo_dist": 4 * cap}
    bufs = {k: api.DeviceBuffer(v) for k, v in spec.items()}
    kitti_ex.stereo_frames_dev(dl.ptr, dr.ptr, 1, w, h, {k: b.ptr for k, b in bufs.items()})
    nl = int(bufs["n_l"].download((1,), np.int32)[0])
    assert nl == host["n_l"][0]
    assert np.array_equal(bufs["kps_l"].download((cap,), api.KP_DTYPE)[:nl], host["kps_l"][0, :nl])
    assert np.array_equal(bufs["desc_r"].download((cap, 32), np.uint8)[:host["n_r"][0]], host["desc_r"][0, :host["n_r"][0]])
    assert np.array_equal(bufs["stereo_idx"].download((cap,), np.int32)[:nl], host["stereo_idx"][0, :nl])
    # extract_batch_dev
    kitti_ex.extract_batch_dev(dl.ptr, 1, w, h, bufs["kps_r"].ptr, bufs["desc_r"].ptr, bufs["n_r"].ptr)
    n = int(bufs["n_r"].download((1,), np.int32)[0])
    assert n == nl and np.array_equal(bufs["desc_r"].download((cap, 32), np.uint8)[:n], host["desc_l"][0, :n])


def _stereo_equal(a, b, frames):
    for f in range(frames):
        nl, nr = a["n_l"][f], a["n_r"][f]
        assert (nl, nr) == (b["n_l"][f], b["n_r"][f])
        for k, n in (("kps_l", nl), ("desc_l", nl), ("kps_r", nr), ("desc_r", nr), ("stereo_idx", nl), ("stereo_dist", nl)):
            assert np.array_equal(a[k][f, :n], b[k][f, :n]), (k, f)


def test_async_queue_with_matchers_on_the_side_stream(gpu):
    """Asynchronous resident calls queue back to back; StereoMatch + tracking of call i run on a side stream beside the
    extraction kernels of call i+1.  Calls that write the SAME output arrays and calls that write different ones must both
    end with exactly what synchronous calls produce -- also when a host-path call or a re-plan follows a pending tail."""
    F, w, h = 3, 640, 240
    seqs = [synth.stereo_sequence(s, F, 4, w, h) for s in (1, 2, 3)]
    cam = api.Camera.make(400.0, 400.0, w / 2.0, h / 2.0, (0, 0, 0, 0), w, h)
    tp = api.TrackParams.make(cam, 0.5, None, 30.0)
    ex = api.ORBextractor(800, 1.2, 5, 20, 7, max_images=2 * F)
    cap = ex.cap
    sync = [ex.stereo_sequence(L, R, tp) for L, R in seqs]
    spec = {"kps_l": (28, api.KP_DTYPE, (F, cap)), "desc_l": (32, np.uint8, (F, cap, 32)), "n_l": (0, np.int32, (F,)),
            "kps_r": (28, api.KP_DTYPE, (F, cap)), "desc_r": (32, np.uint8, (F, cap, 32)), "n_r": (0, np.int32, (F,)),
            "stereo_idx": (4, np.int32, (F, cap)), "stereo_dist": (4, np.int32, (F, cap)), "track_idx": (4, np.int32, (F, cap)),
            "track_dist": (4, np.int32, (F, cap))}
    imgs = [(api.DeviceBuffer(L.nbytes).upload(L), api.DeviceBuffer(R.nbytes).upload(R)) for L, R in seqs]
    sets = [{k: api.DeviceBuffer(max(b * cap * F, 4 * F)) for k, (b, _, _) in spec.items()} for _ in range(3)]

    def fetch(bufs):
        return {k: bufs[k].download(shape, dt) for k, (_, dt, shape) in spec.items()}

    def same(a, b):
        _stereo_equal(a, b, F)
        for f in range(F):
            n = b["n_l"][f]
            assert np.array_equal(a["track_idx"][f, :n], b["track_idx"][f, :n]) and np.array_equal(a["track_dist"][f, :n], b["track_dist"][f, :n])

    ex.set_async(True)
    for rep in range(2):   # different output arrays per call, twice over
        for i in range(3):
            ex.stereo_sequence_dev(imgs[i][0].ptr, imgs[i][1].ptr, F, w, h, {k: b.ptr for k, b in sets[i].items()}, tp)
    ex.wait()
    for i in range(3):
        same(fetch(sets[i]), sync[i])
    for i in (0, 1, 2, 1):  # the same output arrays for every call: the last one must win, untouched by earlier tails
        ex.stereo_sequence_dev(imgs[i][0].ptr, imgs[i][1].ptr, F, w, h, {k: b.ptr for k, b in sets[0].items()}, tp)
    ex.wait()
    same(fetch(sets[0]), sync[1])
    # a tail still pending when a host-path call and a call at another image size arrive
    ex.stereo_sequence_dev(imgs[2][0].ptr, imgs[2][1].ptr, F, w, h, {k: b.ptr for k, b in sets[2].items()}, tp)
    host = ex.stereo_sequence(*seqs[0], tp)
    small = synth.stereo_pair(4, 320, 200)
    ex.stereo_frames(small[0][None], small[1][None])
    ex.wait()
    same(host, sync[0])
    same(fetch(sets[2]), sync[2])
    ex.set_async(False)


def test_resident_batch_split_over_streams(gpu, monkeypatch):
    """A resident batch of >= 48 stereo frames runs as three sub-batches on three streams (SFE_DEV_SPLIT, default 3), also when
    asynchronous calls queue back to back onto the SAME output arrays while the previous call's matchers are still running on
    the side stream: the bytes must equal the unsplit, synchronous run."""
    F, w, h = 48, 320, 200
    def long_sequence(seed):  # three 16-frame sequences back to back
        parts = [synth.stereo_sequence(seed + 10 * k, 16, 4, w, h) for k in range(F // 16)]
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    seqs = [long_sequence(s) for s in (5, 6)]
    cam = api.Camera.make(300.0, 300.0, w / 2.0, h / 2.0, (0, 0, 0, 0), w, h)
    tp = api.TrackParams.make(cam, 0.5, None, 30.0)
    monkeypatch.setenv("SFE_DEV_SPLIT", "1")
    ref_ex = api.ORBextractor(600, 1.2, 4, 20, 7, max_images=2 * F)
    cap = ref_ex.cap
    sync = [ref_ex.stereo_sequence(L, R, tp) for L, R in seqs]
    monkeypatch.delenv("SFE_DEV_SPLIT")
    ex = api.ORBextractor(600, 1.2, 4, 20, 7, max_images=2 * F)
    assert ex.cap == cap
    spec = {"kps_l": (28, api.KP_DTYPE, (F, cap)), "desc_l": (32, np.uint8, (F, cap, 32)), "n_l": (0, np.int32, (F,)),
            "kps_r": (28, api.KP_DTYPE, (F, cap)), "desc_r": (32, np.uint8, (F, cap, 32)), "n_r": (0, np.int32, (F,)),
            "stereo_idx": (4, np.int32, (F, cap)), "stereo_dist": (4, np.int32, (F, cap)), "track_idx": (4, np.int32, (F, cap)),
            "track_dist": (4, np.int32, (F, cap))}
    imgs = [(api.DeviceBuffer(L.nbytes).upload(L), api.DeviceBuffer(R.nbytes).upload(R)) for L, R in seqs]
    bufs = {k: api.DeviceBuffer(max(b * cap * F, 4 * F)) for k, (b, _, _) in spec.items()}
    ex.set_async(True)
    for i in (0, 1, 0, 1, 1, 0):  # the last call must win, untouched by the tails of the earlier ones
        ex.stereo_sequence_dev(imgs[i][0].ptr, imgs[i][1].ptr, F, w, h, {k: b.ptr for k, b in bufs.items()}, tp)
    ex.wait()
    got = {k: bufs[k].download(shape, dt) for k, (_, dt, shape) in spec.items()}
    _stereo_equal(got, sync[0], F)
    for f in range(F):
        n = sync[0]["n_l"][f]
        assert np.array_equal(got["track_idx"][f, :n], sync[0]["track_idx"][f, :n])
        assert np.array_equal(got["track_dist"][f, :n], sync[0]["track_dist"][f, :n])
    ex.set_async(False)
    # the same for the image-only entry point: 96 resident images = three sub-batches of 32
    both = np.concatenate([seqs[0][0], seqs[0][1]])
    dimg = api.DeviceBuffer(both.nbytes).upload(both)
    dk, dd, dn = api.DeviceBuffer(28 * cap * 2 * F), api.DeviceBuffer(32 * cap * 2 * F), api.DeviceBuffer(4 * 2 * F)
    ex.extract_batch_dev(dimg.ptr, 2 * F, w, h, dk.ptr, dd.ptr, dn.ptr)
    n = dn.download((2 * F,), np.int32)
    k, d = dk.download((2 * F, cap), api.KP_DTYPE), dd.download((2 * F, cap, 32), np.uint8)
    for f in range(F):
        for side, key_n, key_k, key_d in ((f, "n_l", "kps_l", "desc_l"), (F + f, "n_r", "kps_r", "desc_r")):
            m = sync[0][key_n][f]
            assert n[side] == m and np.array_equal(k[side, :m], sync[0][key_k][f, :m]) and np.array_equal(d[side, :m], sync[0][key_d][f, :m])


def test_execution_variants_of_the_resident_path_agree(gpu, monkeypatch):
    """Blur forked beside a persistent quadtree or serial, TMA or lane-staged windows in the descriptor kernel, matchers on
    the side stream or not: every switch of DESIGN.md §10 leaves the bytes unchanged."""
    F = 3
    L, R = synth.stereo_sequence(6, F, 4)
    tp = _kitti_track_params()
    base = api.ORBextractor(max_images=2 * F).stereo_sequence(L, R, tp)
    cap, h, w = base["kps_l"].shape[1], *L.shape[1:]
    pitch = api.image_pitch(w)

    def pitched(a):
        out = np.zeros((a.shape[0], h, pitch), np.uint8)
        out[:, :, :w] = a
        return out
    dl, dr = api.DeviceBuffer(F * pitch * h).upload(pitched(L)), api.DeviceBuffer(F * pitch * h).upload(pitched(R))
    spec = {"kps_l": (28, api.KP_DTYPE, (F, cap)), "desc_l": (32, np.uint8, (F, cap, 32)), "n_l": (0, np.int32, (F,)),
            "kps_r": (28, api.KP_DTYPE, (F, cap)), "desc_r": (32, np.uint8, (F, cap, 32)), "n_r": (0, np.int32, (F,)),
            "stereo_idx": (4, np.int32, (F, cap)), "stereo_dist": (4, np.int32, (F, cap)), "track_idx": (4, np.int32, (F, cap)),
            "track_dist": (4, np.int32, (F, cap))}
    for env in ({}, {"SFE_OVERLAP_BLUR": "0"}, {"SFE_OVERLAP_BLUR": "1"}, {"SFE_ORIENT_TMA": "0"}, {"SFE_OVERLAP_TAIL": "0"},
                {"SFE_OCTREE_CTAS": "7"}, {"SFE_NO_TMA": "1", "SFE_OVERLAP_BLUR": "2"}, {"SFE_OCTREE_FF": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ex = api.ORBextractor(max_images=2 * F)
        bufs = {k: api.DeviceBuffer(max(b * cap * F, 4 * F)) for k, (b, _, _) in spec.items()}
        ex.set_async(True)
        for _ in range(2):
            ex.stereo_sequence_dev(dl.ptr, dr.ptr, F, w, h, {k: b.ptr for k, b in bufs.items()}, tp, pitch=pitch)
        ex.wait()
        got = {k: bufs[k].download(shape, dt) for k, (_, dt, shape) in spec.items()}
        _stereo_equal(got, base, F)
        for f in range(F):
            n = base["n_l"][f]
            assert np.array_equal(got["track_idx"][f, :n], base["track_idx"][f, :n]), env
        for k in env:
            monkeypatch.delenv(k)


def test_one_pair_host_call_variants_agree(gpu, oracle, monkeypatch):
    """The reference-shaped call (one stereo pair, host buffers) replays a CUDA graph of kernels launched with programmatic
    stream serialization, starts the quadtree from a directly written depth-d list and stores its results into pinned output
    arrays with one kernel: each of these switched off, pageable instead of pinned buffers, and the oracle must all agree."""
    L, R = synth.stereo_pair(7)
    base = None
    for env in ({}, {"SFE_PDL": "0"}, {"SFE_COPY_KERNEL": "0"}, {"SFE_GRAPHS": "0"}, {"SFE_OCTREE_FF": "0"},
                {"SFE_PDL": "0", "SFE_GRAPHS": "0", "SFE_COPY_KERNEL": "0", "SFE_OCTREE_FF": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ex = api.ORBextractor(max_images=2)
        for pinned in (True, False):
            out = ex.alloc_stereo_out(1, pinned=pinned)
            if pinned:
                pl, pr = api.PinnedArray(L.shape, np.uint8), api.PinnedArray(R.shape, np.uint8)
                pl.array[:], pr.array[:] = L, R
                a, b = pl.array[None], pr.array[None]
            else:
                a, b = L[None], R[None]
            for _ in range(3):  # eager, capture, replay
                got = ex.stereo_frames(a, b, out)
                got = {k: np.array(v, copy=True) for k, v in got.items()}
                if base is None:
                    base = got
                _stereo_equal(got, base, 1)
        for k in env:
            monkeypatch.delenv(k)
    kl, dl = oracle.Extractor().extract(L)
    n = int(base["n_l"][0])
    assert n == len(kl) and np.array_equal(base["desc_l"][0, :n], dl)
    assert np.array_equal(base["kps_l"][0, :n].view(np.uint8).reshape(n, 28), np.ascontiguousarray(kl).view(np.uint8).reshape(n, 28))


def test_pipelined_sub_batches_and_load_paths_agree(gpu, oracle, monkeypatch):
    """The host entry point cuts a batch into sub-batches on several streams, and tiles are fetched either by
    TMA or by plain loads: every combination must give the same bytes, and those of the oracle."""
    seeds = list(range(9))
    L = np.stack([synth.stereo_pair(s)[0] for s in seeds])
    R = np.stack([synth.stereo_pair(s)[1] for s in seeds])
    outs = {}
    for chunks, no_tma in ((1, 0), (3, 0), (9, 0), (2, 1)):
        monkeypatch.setenv("SFE_PIPELINE_CHUNKS", str(chunks))
        monkeypatch.setenv("SFE_NO_TMA", str(no_tma))
        ex = api.ORBextractor(max_images=2 * len(seeds))
        outs[(chunks, no_tma)] = ex.stereo_frames(L, R)
        if chunks == 3:   # odd images too (single-set batch through the same pipeline)
            kps, desc, n = ex.extract_batch(L[:5])
            for i in range(5):
                assert np.array_equal(kps[i, :n[i]], outs[(chunks, no_tma)]["kps_l"][i, :n[i]])
                assert np.array_equal(desc[i, :n[i]], outs[(chunks, no_tma)]["desc_l"][i, :n[i]])
    base = outs[(1, 0)]
    for key, o in outs.items():
        _stereo_equal(o, base, len(seeds))
    ref = oracle.Extractor()
    for f in (0, 8):
        kl, dl = ref.extract(L[f])
        assert np.array_equal(base["kps_l"][f, :len(kl)], kl) and np.array_equal(base["desc_l"][f, :len(kl)], dl)


def test_quadtree_spills_to_global_scratch(gpu, oracle, monkeypatch):
    """Levels with more candidates than the shared-memory arrays hold run the same quadtree code on global scratch."""
    monkeypatch.setenv("SFE_OCTREE_SMEM_CAND", "512")   # every KITTI level has more candidates than this
    L, R = synth.stereo_pair(5)
    ex = api.ORBextractor(max_images=8)            # 2 scratch slots per image slot: 16 for these 16 (image, level) pairs
    out = ex.stereo_frames(L[None], R[None])
    ref = oracle.Extractor()
    for img, kk, dd, nn in ((L, "kps_l", "desc_l", "n_l"), (R, "kps_r", "desc_r", "n_r")):
        rk, rd = ref.extract(img)
        assert out[nn][0] == len(rk)
        assert np.array_equal(out[kk][0, :len(rk)], rk) and np.array_equal(out[dd][0, :len(rk)], rd)


def test_pitched_resident_images_and_async_mode(kitti_ex):
    """Resident images with a TMA-friendly row pitch give the bytes of the host path; in asynchronous mode the
    _dev call returns before the kernels finish and wait() reports."""
    seeds = (1, 2)
    L = np.stack([synth.stereo_pair(s)[0] for s in seeds])
    R = np.stack([synth.stereo_pair(s)[1] for s in seeds])
    host = kitti_ex.stereo_frames(L, R)
    f, h, w = L.shape
    pitch = api.image_pitch(w)
    assert pitch % 16 == 0 and pitch >= w

    def pitched(a):
        out = np.full((f, h, pitch), 255, np.uint8)   # padding must not influence anything
        out[:, :, :w] = a
        return out
    dl, dr = api.DeviceBuffer(f * h * pitch).upload(pitched(L)), api.DeviceBuffer(f * h * pitch).upload(pitched(R))
    cap = kitti_ex.cap
    spec = {"kps_l": 28 * cap * f, "desc_l": 32 * cap * f, "n_l": 4 * f, "kps_r": 28 * cap * f, "desc_r": 32 * cap * f,
            "n_r": 4 * f, "stereo_idx": 4 * cap * f, "stereo_dist": 4 * cap * f}
    for async_mode in (False, True):
        bufs = {k: api.DeviceBuffer(v) for k, v in spec.items()}
        kitti_ex.set_async(async_mode)
        kitti_ex.stereo_frames_dev(dl.ptr, dr.ptr, f, w, h, {k: b.ptr for k, b in bufs.items()}, pitch=pitch)
        if async_mode:
            kitti_ex.wait()
        kitti_ex.set_async(False)
        got = {"n_l": bufs["n_l"].download((f,), np.int32), "n_r": bufs["n_r"].download((f,), np.int32),
               "kps_l": bufs["kps_l"].download((f, cap), api.KP_DTYPE), "kps_r": bufs["kps_r"].download((f, cap), api.KP_DTYPE),
               "desc_l": bufs["desc_l"].download((f, cap, 32), np.uint8), "desc_r": bufs["desc_r"].download((f, cap, 32), np.uint8),
               "stereo_idx": bufs["stereo_idx"].download((f, cap), np.int32),
               "stereo_dist": bufs["stereo_dist"].download((f, cap), np.int32)}
        _stereo_equal(got, host, f)


# ---- matchers -------------------------------------------------------------------------------------
def _mk_kps(xy):
    k = np.zeros(len(xy), api.KP_DTYPE)
    k["x"], k["y"] = xy[:, 0], xy[:, 1]
    return k


def test_stereo_match_random_and_edges(gpu, oracle):
    m = api.Matcher()
    rng = np.random.default_rng(3)
    for n in (1, 31, 300, 2500):
        xyr = np.stack([rng.uniform(0, 1241, n), rng.integers(0, 300, n) * 1.2], 1).astype(np.float32)
        dr = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        sel = rng.integers(0, n, n)
        xyl = xyr[sel] + np.stack([rng.uniform(-5, 110, n), rng.integers(-4, 5, n) * 1.0], 1).astype(np.float32)
        dl = dr[sel].copy()
        flips = rng.integers(0, 60, n)
        for i in range(n):
            for b in rng.integers(0, 256, int(flips[i])):
                dl[i, b >> 3] ^= np.uint8(1 << (b & 7))
        idx, dist = m.StereoMatch(_mk_kps(xyl), dl, _mk_kps(xyr), dr)
        ri, rd = oracle.stereo_match(_mk_kps(xyl), dl, _mk_kps(xyr), dr)
        assert np.array_equal(idx, ri) and np.array_equal(dist, rd), n
    # coordinates far beyond the bucket grid (clamped buckets), wide dx / dy windows, > 2048 right keypoints (chunks)
    for n, span, sp in ((3000, (5000, 3000), api.StereoParams(10.0, 500.0, 0.8)), (700, (300, 2500), api.StereoParams(0.0, 3.0, 0.5)),
                        (500, (4000, 40), api.StereoParams(3.0, 5000.0, 0.9))):
        xyr = np.stack([rng.uniform(-20, span[0], n), rng.integers(-5, span[1], n) * 1.0], 1).astype(np.float32)
        dr = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        sel = rng.integers(0, n, n)
        xyl = xyr[sel] + np.stack([rng.uniform(-5, sp.max_dx * 1.1, n), rng.integers(-12, 13, n) * 1.0], 1).astype(np.float32)
        dl = dr[sel] ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
                        & rng.integers(0, 256, (n, 32), dtype=np.uint8))
        idx, dist = m.StereoMatch(_mk_kps(xyl), dl, _mk_kps(xyr), dr, sp)
        ri, rd = oracle.stereo_match(_mk_kps(xyl), dl, _mk_kps(xyr), dr, sp.y_threshold, sp.max_dx, sp.best12_threshold)
        assert np.array_equal(idx, ri) and np.array_equal(dist, rd), (n, span)
        assert (ri >= 0).sum() > 10
    d = np.zeros((3, 32), np.uint8)
    d[1, 0] = 0xFF
    kl = _mk_kps(np.array([[50.0, 10.0]], np.float32))
    assert m.StereoMatch(kl, d[:1], _mk_kps(np.array([[40.0, 10.0]], np.float32)), d[1:2])[0].tolist() == [0]  # single candidate
    kr = _mk_kps(np.array([[40.0, 10.0], [30.0, 12.0]], np.float32))
    assert m.StereoMatch(kl, d[:1], kr, np.stack([d[1], d[1]]))[0].tolist() == [-1]  # tie for best: rejected
    assert m.StereoMatch(kl, d[:1], _mk_kps(np.array([[50.0, 13.0]], np.float32)), d[1:2])[0].tolist() == [0]
    assert m.StereoMatch(kl, d[:1], _mk_kps(np.array([[50.0, 13.5]], np.float32)), d[1:2])[0].tolist() == [-1]
    assert m.StereoMatch(kl, d[:1], _mk_kps(np.array([[-50.0, 10.0]], np.float32)), d[1:2])[0].tolist() == [0]
    assert m.StereoMatch(kl, d[:1], _mk_kps(np.array([[50.5, 10.0]], np.float32)), d[1:2])[0].tolist() == [-1]
    assert m.StereoMatch(kl, d[:1], kr[:0], d[:0])[0].tolist() == [-1]
    assert m.StereoMatch(kl[:0], d[:0], kr, d[:2])[0].tolist() == []


def test_projection_match_vs_oracle(gpu, oracle, kitti_ex):
    m = api.Matcher()
    L, _ = synth.stereo_pair(0)
    kps, desc = kitti_ex.extract(L)
    xy = np.stack([kps["x"], kps["y"]], 1)
    th = 0.01
    poses = [np.eye(3, 4), np.array([[np.cos(th), 0, np.sin(th), 0.05], [0, 1, 0, -0.02], [-np.sin(th), 0, np.cos(th), 0.2]])]
    for n, seed in ((5000, 1), (60000, 2)):
        xw, mpd = synth.projection_scene(xy, desc, n, seed=seed)
        skip = (np.random.default_rng(seed).uniform(0, 1, n) < 0.05).astype(np.uint8)
        for dcoef in ([0, 0, 0, 0], [-0.05, 0.01, 0.001, -0.002]):
            for rt in poses:
                for radius in (10.0, 50.0, 100.0):
                    if n > 5000 and (radius != 50.0 or rt is poses[1]):
                        continue
                    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
                    ocam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
                    got, gd = m.ProjectionMatch(xw, mpd, skip, rt, cam, kps, desc, radius)
                    ref, rd = oracle.projection_match(xw, mpd, skip, rt, ocam, kps, desc, radius)
                    assert np.array_equal(got, ref), f"n={n} r={radius}: {(got != ref).sum()} keypoints differ"
                    assert np.array_equal(gd, rd)
    assert (ref >= 0).sum() > 100
    # conflict rule: equal distance -> the later query wins; skip mask; behind camera
    kp1 = _mk_kps(np.array([[600.0, 180.0]], np.float32))
    kd1 = np.zeros((1, 32), np.uint8)
    z = 10.0
    X = np.array([[(600.0 - synth.KITTI_CX) / synth.KITTI_FX * z, (180.0 - synth.KITTI_CY) / synth.KITTI_FY * z, z]] * 3)
    mpd = np.zeros((3, 32), np.uint8)
    mpd[0, 0], mpd[1, 0], mpd[2, 1] = 0x03, 0x01, 0x80
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    assert m.ProjectionMatch(X, mpd, None, np.eye(3, 4), cam, kp1, kd1, 50.0)[0].tolist() == [2]
    assert m.ProjectionMatch(X, mpd, np.array([0, 0, 1], np.uint8), np.eye(3, 4), cam, kp1, kd1, 50.0)[0].tolist() == [1]
    Xb = X.copy()
    Xb[:, 2] = -z
    assert m.ProjectionMatch(Xb, mpd, None, np.eye(3, 4), cam, kp1, kd1, 50.0)[0].tolist() == [-1]
    assert m.ProjectionMatch(X[:0], mpd[:0], None, np.eye(3, 4), cam, kp1, kd1, 50.0)[0].tolist() == [-1]


def test_knn2_vs_oracle_and_sharded_merge(gpu, oracle):
    m = api.Matcher()
    db = synth.knn_database(30011, seed=1)
    q, rows = synth.knn_queries(db, 333, seed=2)
    ref = oracle.knn2(q, db)
    full = m.knn2(m.create_db(db), q)
    assert np.array_equal(full, ref)
    # few queries run the row-streaming kernel (thread = row); every lane-count edge, plus duplicated rows (ties)
    db2 = db.copy()
    db2[20000] = db2[77]
    db2[29999] = db2[77]
    dbh = m.create_db(db2)
    for nq in (1, 2, 3, 31, 32, 33, 64, 65, 100, 128, 129):
        qq = q[:nq].copy()
        qq[0] = db2[77]
        got = m.knn2(dbh, qq)
        assert np.array_equal(got, oracle.knn2(qq, db2)), nq
        assert got[0].tolist() == [77, 0, 20000, 0]
    # tiny and degenerate shards
    for rows_n in (0, 1, 2, 255, 257):
        got = m.knn2(m.create_db(db[:rows_n]), q[:5])
        assert np.array_equal(got, oracle.knn2(q[:5], db[:rows_n])), rows_n
    # sharded: per-shard keys -> gathered -> merge == single pass, independent of the shard boundaries
    Q = len(q)
    qd = api.DeviceBuffer(q.nbytes).upload(q)
    for bounds in ([0, 10000, 30011], [0, 1, 7777, 7778, 30011]):
        G = len(bounds) - 1
        keys = api.DeviceBuffer(G * Q * 2 * 8)
        for s in range(G):
            shard = m.create_db(db[bounds[s]:bounds[s + 1]], idx_base=bounds[s])
            m.knn2_dev(shard, qd.ptr, Q, keys.ptr + s * Q * 2 * 8)
        out = api.DeviceBuffer(Q * 16)
        m.knn2_merge_dev(keys.ptr, G, Q, out.ptr)
        assert np.array_equal(out.download((Q, 4), np.int32), ref), bounds


@pytest.mark.parametrize("rows,q", [(1000, 300), (128, 256), (50_001, 512), (50_001, 700), (200_003, 2000), (33, 1300), (3, 257), (70_001, 64), (70_001, 130), (65_536, 255)])
def test_knn2_tensor_core_kernel_equals_popc_kernel_and_oracle(gpu, oracle, monkeypatch, rows, q):
    """From 256 queries on (64 on a large map), the pair distances come from tcgen05.mma kind::i8 on unpacked descriptor bits (sfe_knn_tc.cu);
    SFE_KNN_TC=0 keeps the XOR / POPC kernel.  Both must return the oracle's {idx0, dist0, idx1, dist1}: partial query
    groups, partial row tiles, fewer rows than two, duplicated rows (ties broken by index)."""
    db = synth.knn_database(rows, seed=11)
    qs, _ = synth.knn_queries(db, q, seed=12, hard_fraction=0.2)
    if rows > 300:
        db[rows - 1] = db[7]
        qs[1] = db[7]
    ref = oracle.knn2(qs, db, nthreads=os.cpu_count() or 4)
    # tensor cores with the map's rows kept as ready-made operand tiles (bulk copies; maps of >= 65536 rows searched with >= 256
    # queries), tensor cores with the producer warps unpacking the rows, and the XOR / POPC kernels
    for mode, tiles_gb in (("1", None), ("1", "0"), ("0", None)):
        monkeypatch.setenv("SFE_KNN_TC", mode)
        if tiles_gb is not None:
            monkeypatch.setenv("SFE_KNN_TILES_MAX_GB", tiles_gb)
        m = api.Matcher()
        h = m.create_db(db)
        for _ in range(2):  # the second search finds the tiles already built
            got = m.knn2(h, qs)
            assert np.array_equal(got, ref), f"SFE_KNN_TC={mode} tiles={tiles_gb}: {(got != ref).any(1).sum()} of {q} queries differ"
        if tiles_gb is not None:
            monkeypatch.delenv("SFE_KNN_TILES_MAX_GB")
    if rows > 300:
        assert ref[1].tolist() == [7, 0, rows - 1, 0]


def test_knn2_baseline_config4_full_size_bit_exact(gpu, oracle):
    """BASELINE config 4 exactly as SURVEY §8d states it: 10 M x 32 B map from default_rng(1234), 2000 queries = map rows
    picked by default_rng(5678) with bit flips (a fifth of them flipped hard enough to FAIL the ratio test), every
    {idx0, dist0, idx1, dist1} compared with the threaded C oracle -- and the same over 3 uneven row shards + merge."""
    import os
    m = api.Matcher()
    db = synth.knn_database(10_000_000, seed=1234)
    q, rows = synth.knn_queries(db, 2000, seed=5678, hard_fraction=0.2)
    q[7] = db[123]          # exact duplicates of one row: (0, 123) then the tie partner by index
    db[9_999_999] = db[123]
    ref = oracle.knn2(q, db, nthreads=os.cpu_count() or 8)
    dbh = m.create_db(db)
    out = m.knn2(dbh, q)
    assert np.array_equal(out, ref), f"{(out != ref).any(1).sum()} of 2000 queries differ"
    assert out[7].tolist() == [123, 0, 9_999_999, 0]
    passed = 2 * out[:, 1] < out[:, 3]
    assert 0.6 < passed.mean() < 0.95, "the ratio test must both pass and fail on this workload"
    for nq in (1, 2, 4):    # the streaming kernels (thread = row)
        assert np.array_equal(m.knn2(dbh, q[:nq]), ref[:nq])
    del dbh
    # shards: contiguous, uneven, merged through the packed keys exactly as the multi-GPU path does
    bounds = [0, 3_333_333, 3_333_334, 10_000_000]
    keys = api.DeviceBuffer(3 * 2000 * 2 * 8)
    d_q = api.DeviceBuffer(q.nbytes).upload(q)
    for sidx in range(3):
        lo, hi = bounds[sidx], bounds[sidx + 1]
        sh = m.create_db(db[lo:hi], idx_base=lo)
        m.knn2_dev(sh, d_q.ptr, 2000, keys.ptr + sidx * 2000 * 2 * 8)
        del sh
    d_out = api.DeviceBuffer(2000 * 16)
    m.knn2_merge_dev(keys.ptr, 3, 2000, d_out.ptr)
    assert np.array_equal(d_out.download((2000, 4), np.int32), ref)


def test_projection_match_baseline_config5_full_size_bit_exact(gpu, oracle, kitti_ex):
    """BASELINE config 5 at its stated size: 500 k map points against the seed-0 frame, r = 50, identity pose given as the
    reference's SE3Quat -- every keypoint's (map point, distance) equals the C oracle (grid-accelerated, which
    tests/test_oracle_matchers.py ties to the exhaustive one), plus a real pose with distortion and a skip mask."""
    m = api.Matcher()
    L, _ = synth.stereo_pair(0)
    kps, desc = kitti_ex.extract(L)
    xy = np.stack([kps["x"], kps["y"]], 1)
    xw, mpd = synth.projection_scene(xy, desc, 500_000, seed=99)
    for dcoef, pose, skip in (([0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0], None),
                              ([-0.05, 0.01, 0.001, -0.002], [0.003, -0.004, 0.001, 0.99998650, 0.05, -0.02, 0.2],
                               (np.random.default_rng(5).uniform(0, 1, len(xw)) < 0.05).astype(np.uint8))):
        cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
        ocam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
        qt = np.array(pose, np.float64)
        qt[:4] /= np.linalg.norm(qt[:4])
        got, gd = m.ProjectionMatch(xw, mpd, skip, qt, cam, kps, desc, 50.0)
        ref, rd = oracle.projection_match(xw, mpd, skip, qt, ocam, kps, desc, 50.0, grid=True)
        assert np.array_equal(got, ref), f"{(got != ref).sum()} keypoints differ"
        assert np.array_equal(gd, rd)
        assert (ref >= 0).sum() > 300
    # the matrix form of the identity pose gives the same matches
    got2, _ = m.ProjectionMatch(xw, mpd, None, np.eye(3, 4), api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX,
                                synth.KITTI_CY, [0, 0, 0, 0], 1241, 376), kps, desc, 50.0)
    ref2, _ = oracle.projection_match(xw, mpd, None, np.array([0, 0, 0, 1, 0, 0, 0.0]), oracle.make_camera(
        synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0, 0, 0, 0], 1241, 376), kps, desc, 50.0, grid=True)
    assert np.array_equal(got2, ref2)


def test_projection_match_se3_golden_fixture_of_the_reference(gpu, kitti_ex):
    """tests/golden/golden_proj.npz: the result of the reference's own ProjectionMatch (oracle/_ref) on 3000 map points
    with a distorted camera and a real SE3 pose; keypoints from the seed-0 golden frame"""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_proj.npz"))
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_seed0.npz"))
    m = api.Matcher()
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, z["dist4"], 1241, 376)
    for radius in (50, 10):
        got, _ = m.ProjectionMatch(z["xw"], z["mp_desc"], z["skip"], z["qt"], cam, g["kl"], g["dl"], float(radius))
        assert np.array_equal(got, z[f"to_query_r{radius}"])
    fr = api.Frame(m, g["kl"], g["dl"], cam)
    got, _ = fr.ProjectionMatch(z["xw"], z["mp_desc"], z["skip"], z["qt"], 50.0)
    assert np.array_equal(got, z["to_query_r50"])


# ---- frame glue -----------------------------------------------------------------
def test_resident_frame_glue_vs_oracle(gpu, oracle, kitti_ex):
    """Frame::Frame's post-extraction work, StereoFrame::GetDepth, SearchRadius / SearchNeareast and ProjectionMatch
    against a resident frame: doubles compared bit for bit."""
    m = api.Matcher()
    L, R = synth.stereo_pair(4)
    out = kitti_ex.stereo_frames(L[None], R[None])
    nl, nr = int(out["n_l"][0]), int(out["n_r"][0])
    kps, desc, kps_r, sidx = out["kps_l"][0, :nl], out["desc_l"][0, :nl], out["kps_r"][0, :nr], out["stereo_idx"][0, :nl]
    for dcoef in ([0.0, 0.0, 0.0, 0.0], [-0.05, 0.01, 0.001, -0.002]):
        cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
        ocam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
        f = api.Frame(m, kps, desc, cam)
        nrm = f.normalized()
        ref_n = oracle.normalized_undistort(ocam, kps)
        assert np.array_equal(nrm.view(np.uint64), ref_n.view(np.uint64)), "normalised keypoints differ in some bit"
        xc, valid = f.stereo_depth(kps_r, sidx, 0.54)
        rxc, rvalid = oracle.stereo_depth(ocam, 0.54, kps, ref_n, kps_r, sidx)
        assert np.array_equal(valid, rvalid) and np.array_equal(xc.view(np.uint64), rxc.view(np.uint64))
        assert (valid == 1).sum() == (sidx >= 0).sum() > 100
        ok = (valid == 1) & (kps["octave"] == 0)
        assert np.isclose(np.median(xc[ok, 2]), synth.KITTI_FX * 0.54 / 24.0)   # the synthetic pair has 24 px disparity
        # radius / nearest search
        rng = np.random.default_rng(3)
        uv = np.stack([rng.uniform(-20, 1260, 200), rng.uniform(-20, 396, 200)], 1)
        uv[0] = [kps["x"][5], kps["y"][5]]
        for radius in (10.0, 50.0):
            got = f.SearchRadius(uv, radius)
            for i in range(len(uv)):
                ref, cnt = oracle.search_radius(kps, uv[i, 0], uv[i, 1], radius)
                assert cnt == len(got[i]) and np.array_equal(got[i], ref), (i, radius)
        gi, gd = f.SearchNeareast(uv)
        for i in range(len(uv)):
            ri, rd = oracle.search_nearest(kps, uv[i, 0], uv[i, 1])
            assert gi[i] == ri and gd[i] == rd
        assert gi[0] == 5 and gd[0] == 0.0
        # reprojection error of each keypoint's own stereo point under a slightly wrong pose (ReprojectionFilter::GetOutlier)
        th = 0.004
        Tcw = np.array([[np.cos(th), 0, np.sin(th), 0.03], [0, 1, 0, -0.01], [-np.sin(th), 0, np.cos(th), -0.4], [0, 0, 0, 1.0]])
        xmp = xc.copy()
        xmp[5] = [0.5, 0.2, 0.1]           # lands behind the camera after the -0.4 m shift
        has = (valid == 1).astype(np.uint8)
        has[5] = 1
        err = f.reprojection_error(xmp, has, Tcw)
        ref_err = oracle.reprojection_error(ocam, Tcw, kps, xmp, has)
        assert np.array_equal(err.view(np.uint64), ref_err.view(np.uint64))
        assert np.isinf(err[5]) and (err[has == 0] == -1).all() and 0 < np.median(err[(has == 1) & np.isfinite(err)]) < 30
        # ProjectionMatch against the resident frame == the plain entry point == the oracle
        xy = np.stack([kps["x"], kps["y"]], 1)
        xw, mpd = synth.projection_scene(xy, desc, 20000, seed=8)
        a, ad = f.ProjectionMatch(xw, mpd, None, np.eye(3, 4), 50.0)
        b, bd = m.ProjectionMatch(xw, mpd, None, np.eye(3, 4), cam, kps, desc, 50.0)
        c, cd = oracle.projection_match(xw, mpd, None, np.eye(3, 4), ocam, kps, desc, 50.0)
        assert np.array_equal(a, b) and np.array_equal(ad, bd) and np.array_equal(a, c) and np.array_equal(ad, cd)
    # degenerate frames
    cam = api.Camera.make(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    e = api.Frame(m, kps[:0], desc[:0], cam)
    assert e.normalized().shape == (0, 2) and e.SearchNeareast([[5.0, 5.0]])[0].tolist() == [-1]
    assert len(e.SearchRadius([[5.0, 5.0]], 50.0)[0]) == 0


def test_bow_transform_vs_oracle(gpu, oracle, kitti_ex):
    """Frame::ComputeBoW (src/frame.cpp:419-427): per-feature descent on the GPU against the oracle, BowVector
    assembly against the literal restatement; real ORB descriptors through a synthetic 10^4-word tree."""
    from test_oracle_matchers import _bow_literal
    m = api.Matcher()
    L_img, _ = synth.stereo_pair(1)
    _, desc = kitti_ex.extract(L_img)
    for (k, L, seed) in ((10, 4, 7), (3, 6, 8), (20, 2, 9)):
        parent, is_leaf, ndesc, weight, L = synth.vocabulary(k=k, L=L, seed=seed)
        voc = api.Vocabulary(m, parent, is_leaf, ndesc, weight, L)
        assert voc.words() == int(is_leaf.sum())
        rng = np.random.default_rng(seed)
        feats = np.concatenate([desc, ndesc[rng.integers(1, len(ndesc), 500)]])   # image descriptors + exact node hits (ties)
        for levelsup in (4, 0, 1):
            wid, w, nid = voc.transform_features(feats, levelsup)
            rwid, rw, rnid = oracle.vocab_transform(parent, is_leaf, ndesc, weight, L, feats, levelsup)
            assert np.array_equal(wid, rwid) and np.array_equal(w.view(np.uint64), rw.view(np.uint64)) and np.array_equal(nid, rnid)
        for weighting in range(4):
            for norm in range(3):
                ids, vals = api.bow_assemble(rwid, rw, weighting, norm)
                lit = _bow_literal(rwid, rw, weighting, norm)
                assert ids.tolist() == [a for a, _ in lit]
                assert np.array_equal(vals.view(np.uint64), np.array([b for _, b in lit], np.float64).view(np.uint64))
        (ids, vals), fv = voc.transform(feats, 4)
        assert abs(vals.sum() - 1.0) < 1e-12 and sorted(fv) == list(fv)
        assert sum(len(v) for v in fv.values()) == int((rw > 0).sum())
    assert voc.transform_features(feats[:0])[0].shape == (0,)


def test_copy_probe_and_write_combined_pages(gpu):
    """sfe_copy_probe (the e2e ceiling bench.py prints) and sfe_host_alloc_ex: sane rates, and a write-combined input buffer
    feeds the host entry point with the same results as an ordinary pinned one"""
    up, down = api.copy_probe(0, 8 << 20, 2 << 20, chunks=4, seconds=0.05)
    assert 1.0 < up < 200.0 and 0.5 < down < 200.0
    L, R = synth.stereo_pair(6)
    ex = api.ORBextractor(max_images=2)
    a, b = api.PinnedArray(L.shape, np.uint8), api.PinnedArray(L.shape, np.uint8, write_combined=True)
    a.array[:], b.array[:] = L, L
    k0, d0 = ex.extract(a.array)
    k1, d1 = ex.extract(b.array)
    assert np.array_equal(k0, k1) and np.array_equal(d0, d1) and len(k0) > 1900
