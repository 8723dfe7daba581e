#!/usr/bin/env python3
"""Generate tests/golden/ from oracle/_ref: the REFERENCE'S OWN src/orb_extractor.cpp, src/matcher.cpp and src/camera.cpp
compiled unmodified (oracle/ref_build/Makefile), run with the monotonic heap (address order = creation order, rule T1).

Run in the authoring container (needs /root/reference to build _ref).  The fixtures travel to the GPU box, where
/root/reference never exists.
  golden_seed0.npz          full keypoints/descriptors L+R for synthetic seed 0 + stereo indices
  golden_small{0..3}.npz    keypoints/descriptors of four small / odd configurations
  golden_proj.npz           a ProjectionMatch case (3000 map points, distorted camera, a real SE3 pose) with its result
  golden_hashes.json        SHA-256 of inputs, per-stage outputs and final outputs, seeds 0-7 + small cases + a
                            4-frame stereo sequence with its tracking matches; "generator" records the provenance and
                            "malloc_heap" how far the reference moves under glibc's heap (the address-order tie rule)
StereoFrame::GetDepth (src/frame.cpp:391-409) lives in a translation unit that cannot be compiled here (FLANN/DBoW2);
the sequence section takes the map points from the C oracle's restatement of it and matches them with the reference's
ProjectionMatch.  Distances are not returned by the reference; they are DescriptorDistance of the matched pairs.
"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_c as occ, ref_c as ref
from slam_toolkit_b200 import synth

ref.set_heap_mode(ref.HEAP_MONOTONIC)
IDENT = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(ex, img):
    k, d = ex.extract(img)
    nl = ex.nlevels
    pyr = [ex.level(l) for l in range(nl)]
    cands = [ex.candidates(l) for l in range(nl)]
    quota = ex.tables()["per_level"]
    dist = []
    for l in range(nl):
        w, h = ex.level_size(l)
        dist.append(ex.distribute(cands[l], 16, w - 16, 16, h - 16, int(quota[l]), l) if len(cands[l]) else np.zeros((0, 3), np.float32))
    ref.set_heap_mode(ref.HEAP_MONOTONIC)
    blur = [ex.blur(l) for l in range(nl)]
    rec = {"input": sha(img), "n": int(len(k)), "kps": sha(k), "desc": sha(d),
           "pyramid": [sha(p) for p in pyr], "blur": [sha(b) if b is not None else None for b in blur],
           "cands": [sha(c) for c in cands], "ncands": [int(len(c)) for c in cands],
           "dist": [sha(c) for c in dist], "ndist": [int(len(c)) for c in dist]}
    assert sum(rec["ndist"]) == rec["n"]
    return k, d, rec


def stereo(kl, dl, kr, dr, cam):
    si = ref.stereo_match(kl, dl, kr, dr, cam)
    sd = np.array([occ.hamming256(dl[i], dr[j]) if j >= 0 else -1 for i, j in enumerate(si)], np.int32)
    return si, sd


def track(cam, prev, kl, dl):
    """ProjectionMatch of the previous frame's stereo points (GetDepth) into the current frame, identity prior, r = 50"""
    pkl, pdl, pkr, psi = prev
    xc, valid = occ.stereo_depth(cam, synth.KITTI_BASELINE, pkl, occ.normalized_undistort(cam, pkl), pkr, psi)
    assert np.array_equal(occ.normalized_undistort(cam, pkl), ref.normalized_undistort(cam, pkl))
    sel = np.nonzero(valid == 1)[0]
    to_q = ref.projection_match(xc[sel], pdl[sel], None, IDENT, cam, kl, dl, 50.0)
    ti = np.where(to_q >= 0, sel[np.maximum(to_q, 0)], -1).astype(np.int32)
    td = np.array([occ.hamming256(pdl[i], dl[j]) if i >= 0 else -1 for j, i in enumerate(ti)], np.int32)
    return ti, td


def sequence_section(cam):
    ex = ref.Extractor(2000, 1.2, 8, 20, 7)
    L, R = synth.stereo_sequence(5, 4, 4)
    rec = {"inputs": [sha(L), sha(R)], "frames": []}
    prev = None
    for f in range(4):
        kl, dl, _ = run(ex, L[f]); kr, dr, _ = run(ex, R[f])
        si, sd = stereo(kl, dl, kr, dr, cam)
        fr = {"kps_l": sha(kl), "desc_l": sha(dl), "stereo_idx": sha(si)}
        if prev is not None:
            ti, td = track(cam, prev, kl, dl)
            fr.update({"track_idx": sha(ti), "track_dist": sha(td), "n_tracked": int((ti >= 0).sum())})
        rec["frames"].append(fr)
        prev = (kl, dl, kr, si)
        print("sequence frame", f, len(kl), fr.get("n_tracked"), flush=True)
    return rec


def projection_case(kl, dl):
    """3000 map points around the seed-0 frame: half carry a frame descriptor with 3 flipped bits, 5 % are already in the
    frame (skip), every 17th lies behind the camera; distorted camera, a small rotation + translation."""
    rng = np.random.default_rng(77)
    n = 3000
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
    skip = (rng.random(n) < 0.05).astype(np.uint8)
    cam = occ.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [-0.1, 0.02, 0.001, -0.0005],
                          synth.KITTI_W, synth.KITTI_H)
    qt_in = np.array([0.01, -0.02, 0.005, 0.9997, 0.1, -0.05, 0.2])
    _, q = ref.se3_apply(qt_in, xw[:1])            # the unit quaternion the SE3Quat holds
    qt = np.concatenate([q, qt_in[4:]])
    res = {}
    for radius in (50.0, 10.0):
        to_q = ref.projection_match(xw, md, skip, qt, cam, kl, dl, radius)
        res[f"to_query_r{int(radius)}"] = to_q
        print("projection r", radius, int((to_q >= 0).sum()), flush=True)
    xc, _ = ref.se3_apply(qt, xw)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/golden_proj.npz"), xw=xw, mp_desc=md, skip=skip, qt=qt,
                        dist4=np.array([-0.1, 0.02, 0.001, -0.0005]), xc=xc, **res)


def malloc_heap_stats(seeds):
    """How far the reference's own result moves under glibc malloc (heap-address tie rule, src/orb_extractor.cpp:684)"""
    ex = ref.Extractor()
    out = {}
    for seed in seeds:
        L, _ = synth.stereo_pair(seed)
        ref.set_heap_mode(ref.HEAP_MONOTONIC)
        k0, d0 = ex.extract(L)
        ref.set_heap_mode(ref.HEAP_MALLOC)
        k1, d1 = ex.extract(L)
        ref.set_heap_mode(ref.HEAP_MONOTONIC)
        a = set(map(bytes, np.concatenate([k0.view(np.uint8).reshape(len(k0), 28), d0], 1)))
        b = set(map(bytes, np.concatenate([k1.view(np.uint8).reshape(len(k1), 28), d1], 1)))
        out[str(seed)] = {"n_monotonic": len(k0), "n_malloc": len(k1), "only_monotonic": len(a - b), "only_malloc": len(b - a)}
    return out


cam = occ.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0, 0, 0, 0], synth.KITTI_W, synth.KITTI_H)
src_sha = open(os.path.join(ROOT, "oracle/_ref/sources.sha256")).read().split()
out = {"generator": {"what": "oracle/_ref: geonuklee/slam-toolkit src/orb_extractor.cpp + src/matcher.cpp + src/camera.cpp compiled "
                             "unmodified against stand-in third-party headers (oracle/ref_build), monotonic heap",
                     "sources_sha256": dict(zip([os.path.basename(p) for p in src_sha[1::2]], src_sha[0::2]))},
       "kitti": {}, "small": {}}
ex = ref.Extractor(2000, 1.2, 8, 20, 7)
for seed in range(8):
    L, R = synth.stereo_pair(seed)
    kl, dl, recl = run(ex, L); kr, dr, recr = run(ex, R)
    si, sd = stereo(kl, dl, kr, dr, cam)
    out["kitti"][str(seed)] = {"L": recl, "R": recr, "stereo_idx": sha(si), "stereo_dist": sha(sd),
                               "n_stereo": int((si >= 0).sum())}
    if seed == 0:
        np.savez_compressed(os.path.join(ROOT, "tests/golden/golden_seed0.npz"), kl=kl, dl=dl, kr=kr, dr=dr,
                            stereo_idx=si, stereo_dist=sd)
        projection_case(kl, dl)
    print("seed", seed, recl["n"], recr["n"], out["kitti"][str(seed)]["n_stereo"], flush=True)
# small / odd configurations: (w, h, nfeatures, scale, nlevels, ini, min)
small = [(320, 240, 500, 1.2, 4, 20, 7), (161, 131, 300, 1.5, 3, 20, 7), (640, 200, 1000, 1.2, 8, 30, 10),
         (97, 95, 50, 1.2, 2, 20, 7)]
for i, (w, h, nf, sf, nl, it, mt) in enumerate(small):
    e = ref.Extractor(nf, sf, nl, it, mt)
    img, _ = synth.stereo_pair(100 + i, w, h)
    k, d, rec = run(e, img)
    rec["params"] = [w, h, nf, sf, nl, it, mt]
    out["small"][str(i)] = rec
    np.savez_compressed(os.path.join(ROOT, f"tests/golden/golden_small{i}.npz"), k=k, d=d)
    print("small", i, rec["n"], rec["ncands"], flush=True)
out["sequence"] = sequence_section(cam)
out["malloc_heap"] = malloc_heap_stats(range(8))
print("malloc heap:", out["malloc_heap"], flush=True)
json.dump(out, open(os.path.join(ROOT, "tests/golden/golden_hashes.json"), "w"), indent=1)
