#!/usr/bin/env python3
"""Generate tests/golden/ from the cv2-backed oracle (oracle/oracle_cv2.py).

Run in the build container (needs cv2 4.13).  The fixtures travel to the GPU box,
where cv2 may or may not exist and /root/reference never does.
  golden_seed0.npz          full keypoints/descriptors L+R for synthetic seed 0 + stereo indices
  golden_hashes.json        SHA-256 of inputs, per-stage outputs and final outputs, seeds 0-7 + small cases + a
                            4-frame stereo sequence with its tracking matches (--sequence-only refreshes that part)
"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "slam-toolkit_b200"))
import synth, oracle_cv2 as oc, oracle_c as occ
import cv2
cv2.setNumThreads(1)
try:
    cv2.ipp.setUseIPP(False)
except Exception:
    pass

def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

def kp_struct(k6):
    out = np.zeros(len(k6), occ.KP_DTYPE)
    out["x"], out["y"], out["size"], out["angle"], out["response"] = k6[:, 0], k6[:, 1], k6[:, 2], k6[:, 3], k6[:, 4]
    out["octave"] = k6[:, 5].astype(np.int32); out["class_id"] = -1
    return out

def run(ex, img):
    st = {}
    k6, d = ex.extract(img, st)
    rec = {"input": sha(img), "n": int(len(k6)), "kps": sha(kp_struct(k6)), "desc": sha(d),
           "pyramid": [sha(p) for p in st["pyramid"]], "blur": [sha(b) for b in st["blur"]],
           "cands": [sha(c) for c in st["cands"]], "ncands": [int(len(c)) for c in st["cands"]],
           "dist": [sha(c) for c in st["dist"]], "ndist": [int(len(c)) for c in st["dist"]]}
    return kp_struct(k6), d, rec

def sequence_section():
    """A 4-frame stereo sequence (scene 5, the camera slides 4 px per frame): extraction by the cv2 restatement, StereoMatch
    and the tracking step (GetDepth + ProjectionMatch, r = 50, identity prior) by the C oracle -- like the matcher they
    have no cv2 primitive; tests/test_oracle_matchers.py holds their literal restatement."""
    ex = oc.ExtractorCv2(2000, 1.2, 8, 20, 7)
    L, R = synth.stereo_sequence(5, 4, 4)
    cam = occ.make_camera(synth.KITTI_FX, synth.KITTI_FY, synth.KITTI_CX, synth.KITTI_CY, [0, 0, 0, 0], synth.KITTI_W, synth.KITTI_H)
    rec = {"inputs": [sha(L), sha(R)], "frames": []}
    prev = None
    for f in range(4):
        kl, dl, _ = run(ex, L[f]); kr, dr, _ = run(ex, R[f])
        si, sd = occ.stereo_match(kl, dl, kr, dr)
        fr = {"kps_l": sha(kl), "desc_l": sha(dl), "stereo_idx": sha(si)}
        if prev is not None:
            ti, td = occ.track_pair(cam, synth.KITTI_BASELINE, np.eye(4), 50.0, prev[0], prev[1], prev[2], prev[3], kl, dl)
            fr.update({"track_idx": sha(ti), "track_dist": sha(td), "n_tracked": int((ti >= 0).sum())})
        rec["frames"].append(fr)
        prev = (kl, dl, kr, si)
        print("sequence frame", f, len(kl), fr.get("n_tracked"), flush=True)
    return rec


if "--sequence-only" in sys.argv:   # add / refresh the sequence section of the committed file
    path = os.path.join(ROOT, "tests/golden/golden_hashes.json")
    out = json.load(open(path))
    out["sequence"] = sequence_section()
    json.dump(out, open(path, "w"), indent=1)
    sys.exit(0)

out = {"cv2": cv2.__version__, "kitti": {}, "small": {}}
ex = oc.ExtractorCv2(2000, 1.2, 8, 20, 7)
for seed in range(8):
    L, R = synth.stereo_pair(seed)
    kl, dl, recl = run(ex, L); kr, dr, recr = run(ex, R)
    si, sd = occ.stereo_match(kl, dl, kr, dr)   # matcher has no cv2 primitive: C oracle is its restatement
    out["kitti"][str(seed)] = {"L": recl, "R": recr, "stereo_idx": sha(si), "stereo_dist": sha(sd),
                               "n_stereo": int((si >= 0).sum())}
    if seed == 0:
        np.savez_compressed(os.path.join(ROOT, "tests/golden/golden_seed0.npz"), kl=kl, dl=dl, kr=kr, dr=dr,
                            stereo_idx=si, stereo_dist=sd)
    print("seed", seed, recl["n"], recr["n"], out["kitti"][str(seed)]["n_stereo"], flush=True)
# small / odd configurations: (w, h, nfeatures, scale, nlevels, ini, min)
small = [(320, 240, 500, 1.2, 4, 20, 7), (161, 131, 300, 1.5, 3, 20, 7), (640, 200, 1000, 1.2, 8, 30, 10),
         (97, 95, 50, 1.2, 2, 20, 7)]
for i, (w, h, nf, sf, nl, it, mt) in enumerate(small):
    e = oc.ExtractorCv2(nf, sf, nl, it, mt)
    img, _ = synth.stereo_pair(100 + i, w, h)
    k, d, rec = run(e, img)
    rec["params"] = [w, h, nf, sf, nl, it, mt]
    out["small"][str(i)] = rec
    np.savez_compressed(os.path.join(ROOT, f"tests/golden/golden_small{i}.npz"), k=k, d=d)
    print("small", i, rec["n"], rec["ncands"], flush=True)
out["sequence"] = sequence_section()
json.dump(out, open(os.path.join(ROOT, "tests/golden/golden_hashes.json"), "w"), indent=1)
