#!/usr/bin/env python3
"""Generate tests/golden/ from the cv2-backed oracle (oracle/oracle_cv2.py).

Run in the build container (needs cv2 4.13).  The fixtures travel to the GPU box,
where cv2 may or may not exist and /root/reference never does.
  golden_seed0.npz          full keypoints/descriptors L+R for synthetic seed 0 + stereo indices
  golden_hashes.json        SHA-256 of inputs, per-stage outputs and final outputs, seeds 0-7 + small cases
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
json.dump(out, open(os.path.join(ROOT, "tests/golden/golden_hashes.json"), "w"), indent=1)
