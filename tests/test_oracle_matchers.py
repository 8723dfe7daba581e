"""Matcher half of the oracle against a slow, literal Python restatement of
reference src/matcher.cpp:54-209 (buckets and std::set order included)."""
import numpy as np
import pytest

from slam_toolkit_b200 import synth


def _ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _stereo_literal(kl, dl, kr, dr):
    rows = {}
    for j in range(len(kr)):
        rows.setdefault(int(np.float64(kr["y"][j]) / 10.0), set()).add(j)
    out = np.full(len(kl), -1, np.int32)
    for i in range(len(kl)):
        y = np.float64(kl["y"][i])
        inl = set()
        for r in (int(y / 10.0), int(y / 10.0 - 1), int(y / 10.0 + 1)):
            inl |= rows.get(r, set())
        d0 = d1 = 999999999.0
        c0 = -1
        for j in sorted(inl):
            dx = np.float64(np.float32(kl["x"][i] - kr["x"][j]))
            dy = np.float64(np.float32(kl["y"][i] - kr["y"][j]))
            if abs(dy) > 3.0 or dx < 0.0 or dx > 100.0:
                continue
            d = float(_ham(dl[i], dr[j]))
            if d < d0:
                d1, d0, c0 = d0, d, j
            elif d < d1:
                d1 = d
        if c0 >= 0 and d0 < d1 * 0.5:
            out[i] = c0
    return out


def _mk_kps(oracle, xy):
    k = np.zeros(len(xy), oracle.KP_DTYPE)
    k["x"], k["y"] = xy[:, 0], xy[:, 1]
    return k


def test_stereo_vs_literal(oracle):
    rng = np.random.default_rng(3)
    n = 300
    xyr = np.stack([rng.uniform(0, 400, n), rng.integers(0, 60, n) * 1.2], 1).astype(np.float32)
    dr = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sel = rng.integers(0, n, n)
    xyl = xyr[sel] + np.stack([rng.uniform(-5, 110, n), rng.integers(-4, 5, n) * 1.0], 1).astype(np.float32)
    dl = dr[sel].copy()
    for i in range(n):
        for b in rng.integers(0, 256, int(rng.integers(0, 60))):
            dl[i, b >> 3] ^= np.uint8(1 << (b & 7))
    kl, kr = _mk_kps(oracle, xyl), _mk_kps(oracle, xyr)
    idx, dist = oracle.stereo_match(kl, dl, kr, dr)
    assert np.array_equal(idx, _stereo_literal(kl, dl, kr, dr))
    assert (idx >= 0).sum() > 20
    ok = idx >= 0
    assert all(dist[i] == _ham(dl[i], dr[idx[i]]) for i in np.nonzero(ok)[0])


def test_stereo_edge_rules(oracle):
    d = np.zeros((3, 32), np.uint8)
    d[1, 0] = 0xFF  # 8 bits from d[0]
    kl = _mk_kps(oracle, np.array([[50.0, 10.0]], np.float32))
    # single candidate: accepted (dist1 stays 999999999)
    idx, _ = oracle.stereo_match(kl, d[:1], _mk_kps(oracle, np.array([[40.0, 10.0]], np.float32)), d[1:2])
    assert idx.tolist() == [0]
    # two candidates tying for best: rejected
    kr = _mk_kps(oracle, np.array([[40.0, 10.0], [30.0, 12.0]], np.float32))
    idx, _ = oracle.stereo_match(kl, d[:1], kr, np.stack([d[1], d[1]]))
    assert idx.tolist() == [-1]
    # filters: dy == 3 passes, dy > 3 fails, dx == 100 passes, dx < 0 fails
    kr = _mk_kps(oracle, np.array([[50.0, 13.0]], np.float32))
    assert oracle.stereo_match(kl, d[:1], kr, d[1:2])[0].tolist() == [0]
    kr = _mk_kps(oracle, np.array([[50.0, 13.5]], np.float32))
    assert oracle.stereo_match(kl, d[:1], kr, d[1:2])[0].tolist() == [-1]
    kr = _mk_kps(oracle, np.array([[-50.0, 10.0]], np.float32))
    assert oracle.stereo_match(kl, d[:1], kr, d[1:2])[0].tolist() == [0]
    kr = _mk_kps(oracle, np.array([[50.5, 10.0]], np.float32))
    assert oracle.stereo_match(kl, d[:1], kr, d[1:2])[0].tolist() == [-1]
    # empty sides
    assert oracle.stereo_match(kl, d[:1], kr[:0], d[:0])[0].tolist() == [-1]
    assert oracle.stereo_match(kl[:0], d[:0], kr, d[1:2])[0].tolist() == []


def _proj_literal(xw, mpd, skip, rt, cam, kps, kpd, radius):
    m = len(kps)
    match, dist = {}, {}
    R = np.asarray(rt, np.float64).reshape(3, 4)
    for i in range(len(xw)):
        if skip is not None and skip[i]:
            continue
        X = xw[i]
        xc = [((R[r, 0] * X[0] + R[r, 1] * X[1]) + R[r, 2] * X[2]) + R[r, 3] for r in range(3)]
        if xc[2] < 0:
            continue
        x, y = xc[0] / xc[2], xc[1] / xc[2]
        r2 = x * x + y * y
        r4 = r2 * r2
        a1, a2, a3 = 2. * x * y, r2 + 2. * x * x, r2 + 2. * y * y
        cd = 1. + cam.d[0] * r2 + cam.d[1] * r4
        xd = x * cd + cam.d[2] * a1 + cam.d[3] * a2
        yd = y * cd + cam.d[2] * a3 + cam.d[3] * a1
        u, v = cam.fx * xd + cam.cx, cam.fy * yd + cam.cy
        if u < 0 or v < 0 or u > cam.width or v > cam.height:
            continue
        d0 = d1 = 999999999.0
        c0 = -1
        for j in range(m):
            dd = (u - float(kps["x"][j])) ** 2 + (v - float(kps["y"][j])) ** 2
            if not dd < radius * radius:
                continue
            d = float(_ham(mpd[i], kpd[j]))
            if d < d0:
                d1, d0, c0 = d0, d, j
            elif d < d1:
                d1 = d
        if c0 < 0 or not d0 < d1 * 0.5:
            continue
        if c0 in match and dist[c0] < d0:
            continue
        match[c0], dist[c0] = i, d0
    out = np.full(m, -1, np.int32)
    for k, v in match.items():
        out[k] = v
    return out


def test_projection_vs_literal(oracle):
    rng = np.random.default_rng(5)
    m = 150
    kxy = np.stack([rng.uniform(0, 1241, m), rng.uniform(0, 376, m)], 1).astype(np.float32)
    kps = _mk_kps(oracle, kxy)
    kpd = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    xw, mpd = synth.projection_scene(kxy, kpd, 600, seed=1)
    skip = (rng.uniform(0, 1, 600) < 0.1).astype(np.uint8)
    th = 0.05
    rt = np.array([[np.cos(th), 0, np.sin(th), 0.1], [0, 1, 0, -0.05], [-np.sin(th), 0, np.cos(th), 0.3]])
    for dcoef in ([0, 0, 0, 0], [-0.05, 0.01, 0.001, -0.002]):
        cam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, dcoef, 1241, 376)
        for radius in (10.0, 50.0):
            got, dist = oracle.projection_match(xw, mpd, skip, rt, cam, kps, kpd, radius)
            ref = _proj_literal(xw, mpd, skip, rt, cam, kps, kpd, radius)
            assert np.array_equal(got, ref)
    assert (got >= 0).sum() > 5


def test_projection_conflict_later_wins(oracle):
    kps = _mk_kps(oracle, np.array([[600.0, 180.0]], np.float32))
    kpd = np.zeros((1, 32), np.uint8)
    cam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    z = 10.0
    X = np.array([[(600.0 - synth.KITTI_CX) / synth.KITTI_FX * z, (180.0 - synth.KITTI_CY) / synth.KITTI_FY * z, z]] * 3)
    mpd = np.zeros((3, 32), np.uint8)
    mpd[0, 0] = 0x03  # dist 2
    mpd[1, 0] = 0x01  # dist 1
    mpd[2, 1] = 0x80  # dist 1, later -> wins the tie
    rt = np.eye(3, 4)
    to_q, dist = oracle.projection_match(X, mpd, None, rt, cam, kps, kpd, 50.0)
    assert to_q.tolist() == [2] and dist.tolist() == [1]
    to_q, _ = oracle.projection_match(X, mpd, np.array([0, 0, 1], np.uint8), rt, cam, kps, kpd, 50.0)
    assert to_q.tolist() == [1]
    # behind the camera / outside the image: no match
    Xb = X.copy(); Xb[:, 2] = -z
    assert oracle.projection_match(Xb, mpd, None, rt, cam, kps, kpd, 50.0)[0].tolist() == [-1]


def test_knn2_and_hamming(oracle):
    db = synth.knn_database(5000, seed=1)
    q, rows = synth.knn_queries(db, 40, seed=2)
    out = oracle.knn2(q, db)
    bits_db = np.unpackbits(db, axis=1).astype(np.int16)
    for i in range(40):
        d = np.abs(bits_db - np.unpackbits(q[i]).astype(np.int16)).sum(1)
        order = np.lexsort((np.arange(len(d)), d))
        assert out[i].tolist() == [order[0], d[order[0]], order[1], d[order[1]]]
        assert oracle.hamming256(q[i], db[rows[i]]) == d[rows[i]]
    # sharded evaluation + lexicographic merge == one pass (exactness of the multi-GPU scheme)
    a = oracle.knn2(q, db[:2000], 0)
    b = oracle.knn2(q, db[2000:], 2000)
    for i in range(40):
        c = sorted([(a[i, 1], a[i, 0]), (a[i, 3], a[i, 2]), (b[i, 1], b[i, 0]), (b[i, 3], b[i, 2])])
        assert [c[0][1], c[0][0], c[1][1], c[1][0]] == out[i].tolist()


def test_frame_glue_oracle_properties(oracle):
    """orc_normalized_undistort inverts Distort (the reference's 5-step fixed point, src/camera.cpp:95-109) and the
    searches agree with a numpy restatement."""
    import numpy as np
    from slam_toolkit_b200 import synth
    rng = np.random.default_rng(0)
    k = np.zeros(500, oracle.KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, 1241, 500).astype(np.float32), rng.uniform(0, 376, 500).astype(np.float32)
    d = [-0.05, 0.01, 0.001, -0.002]
    cam = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, d, 1241, 376)
    n = oracle.normalized_undistort(cam, k)
    x, y = n[:, 0], n[:, 1]
    r2 = x * x + y * y
    cd = 1 + d[0] * r2 + d[1] * r2 * r2
    xd = x * cd + d[2] * 2 * x * y + d[3] * (r2 + 2 * x * x)
    yd = y * cd + d[2] * (r2 + 2 * y * y) + d[3] * 2 * x * y
    # five iterations only (as in the reference): converged to well under a thousandth of a pixel at the image corners
    assert np.abs(xd * synth.KITTI_FX + synth.KITTI_CX - k["x"]).max() < 1e-3
    assert np.abs(yd * synth.KITTI_FY + synth.KITTI_CY - k["y"]).max() < 1e-3
    cam0 = oracle.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0] * 4, 1241, 376)
    n0 = oracle.normalized_undistort(cam0, k)
    assert np.abs(n0[:, 0] - (k["x"].astype(np.float64) - synth.KITTI_CX) / synth.KITTI_FX).max() < 1e-15
    for u, v, r in ((600.0, 180.0, 50.0), (0.0, 0.0, 100.0), (float(k["x"][3]), float(k["y"][3]), 1.0)):
        d2 = (u - k["x"].astype(np.float64)) ** 2 + (v - k["y"].astype(np.float64)) ** 2
        idx, cnt = oracle.search_radius(k, u, v, r)
        assert np.array_equal(idx, np.nonzero(d2 < r * r)[0]) and cnt == len(idx)
        i, dd = oracle.search_nearest(k, u, v)
        assert i == int(np.argmin(d2)) and dd == d2.min()
    xc, valid = oracle.stereo_depth(cam0, 0.5, k[:3], n0[:3], k[3:6], np.array([0, -1, 2], np.int32))
    assert valid.tolist()[1] == 0 and valid[0] in (1, 2)


def _bow_literal(wid, w, weighting, norm):
    """BowVector::addWeight / addIfNotExist + normalize, literally (thirdparty/DBoW2/DBoW2/BowVector.cpp:34-84,
    TemplatedVocabulary.h:1127-1194), with Python floats (IEEE doubles)."""
    import math
    v = {}
    for i in range(len(wid)):
        if not w[i] > 0:
            continue
        k = int(wid[i])
        if weighting in (0, 1):
            v[k] = v[k] + float(w[i]) if k in v else float(w[i])
        elif k not in v:
            v[k] = float(w[i])
    items = sorted(v.items())
    if weighting in (0, 1) and items and norm == 0:
        nd = float(len(items))
        items = [(k, x / nd) for k, x in items]
    if norm:
        s = 0.0
        for _, x in items:
            s = s + (abs(x) if norm == 1 else x * x)
        if norm == 2:
            s = math.sqrt(s)
        if s > 0.0:
            items = [(k, x / s) for k, x in items]
    return items


def test_vocab_transform_oracle_vs_literal(oracle):
    """orc_vocab_transform against a literal Python walk of TemplatedVocabulary::transform (:1218-1259)."""
    import numpy as np
    from slam_toolkit_b200 import synth
    parent, is_leaf, desc, weight, L = synth.vocabulary(k=5, L=3, seed=3, early_leaf=0.1)
    children = {}
    words, wid_of = 0, {}
    for i in range(1, len(parent)):
        children.setdefault(int(parent[i]), []).append(i)
        if is_leaf[i]:
            wid_of[i] = words
            words += 1
    rng = np.random.default_rng(0)
    feats = desc[rng.integers(1, len(desc), 200)].copy()
    for f in feats:
        for b in rng.integers(0, 256, 10):
            f[b >> 3] ^= np.uint8(1 << (b & 7))
    for levelsup in (0, 1, 2, 5):
        wid, w, nid = oracle.vocab_transform(parent, is_leaf, desc, weight, L, feats, levelsup)
        for f in range(len(feats)):
            final, level, n_at = 0, 0, 0
            while True:
                level += 1
                ch = children[final]
                d = [int(np.unpackbits(feats[f] ^ desc[c]).sum()) for c in ch]
                final = ch[int(np.argmin(d))]          # argmin = first minimum = the reference's strict <
                if level == L - levelsup:
                    n_at = final
                if final not in children:
                    break
            assert (wid[f], w[f], nid[f]) == (wid_of.get(final, 0), weight[final], n_at)


def test_grid_projection_match_equals_brute_force_and_track_pair_composes(oracle):
    oc = oracle
    """orc_projection_match_grid (the CPU baseline's stand-in for the FLANN kd-tree) visits the same candidate set as the
    brute-force oracle; orc_track_pair = NormalizedUndistort + GetDepth + ProjectionMatch with the no-depth points skipped."""
    rng = np.random.default_rng(11)
    cam = oc.make_camera(718.856, 718.856, 607.1928, 185.2157, [0.01, -0.002, 0.0005, -0.0003], 1241, 376)
    for trial in range(6):
        m, n = int(rng.integers(1, 600)), int(rng.integers(1, 900))
        kps = np.zeros(m, oc.KP_DTYPE)
        kps["x"] = rng.uniform(-5, 1250, m).astype(np.float32)   # a few keypoints outside the image: clamped cells
        kps["y"] = rng.uniform(-5, 380, m).astype(np.float32)
        desc = rng.integers(0, 256, (m, 32), dtype=np.uint8)
        z = rng.uniform(-1, 60, n)
        xw = np.stack([(rng.uniform(-50, 1300, n) - 607.1928) / 718.856 * z, (rng.uniform(-50, 420, n) - 185.2157) / 718.856 * z, z], 1)
        mpd = desc[rng.integers(0, m, n)] ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
                                             & rng.integers(0, 256, (n, 32), dtype=np.uint8))
        skip = (rng.uniform(size=n) < 0.1).astype(np.uint8)
        rt = np.eye(4)[:3]
        for radius in (0.0, 7.5, 50.0, 400.0):
            a = oc.projection_match(xw, mpd, skip, rt, cam, kps, desc, radius)
            b = oc.projection_match(xw, mpd, skip, rt, cam, kps, desc, radius, grid=True)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (trial, radius)
    # track_pair against its parts
    n_prev, n_cur = 500, 450
    kl = np.zeros(n_prev, oc.KP_DTYPE)
    kl["x"], kl["y"] = rng.uniform(30, 1200, n_prev).astype(np.float32), rng.uniform(30, 350, n_prev).astype(np.float32)
    kr = kl.copy()
    kr["x"] -= rng.uniform(-2, 60, n_prev).astype(np.float32)     # some negative disparities: skipped
    sidx = np.where(rng.uniform(size=n_prev) < 0.7, rng.permutation(n_prev), -1).astype(np.int32)
    dl = rng.integers(0, 256, (n_prev, 32), dtype=np.uint8)
    cur = np.zeros(n_cur, oc.KP_DTYPE)
    src = rng.integers(0, n_prev, n_cur)
    cur["x"], cur["y"] = kl["x"][src] + rng.uniform(-6, 6, n_cur).astype(np.float32), kl["y"][src] + rng.uniform(-6, 6, n_cur).astype(np.float32)
    dc = dl[src] ^ (1 << rng.integers(0, 8, (n_cur, 32))).astype(np.uint8) * (rng.uniform(size=(n_cur, 32)) < 0.1)
    dc = dc.astype(np.uint8)
    nrm = oc.normalized_undistort(cam, kl)
    xc, valid = oc.stereo_depth(cam, 0.537, kl, nrm, kr, sidx)
    ref = oc.projection_match(xc, dl, (valid != 1).astype(np.uint8), np.eye(4)[:3], cam, cur, dc, 30.0)
    for grid in (False, True):
        got = oc.track_pair(cam, 0.537, np.eye(4), 30.0, kl, dl, kr, sidx, cur, dc, grid=grid)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    assert (ref[0] >= 0).sum() > 20
