"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/orb_oracle.c).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("d", C.c_double * 4), ("width", C.c_int), ("height", C.c_int)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "orb_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, ci, cf, cd = C.c_void_p, C.c_int, C.c_float, C.c_double
        L.orc_extractor_create.restype = vp
        L.orc_extractor_create.argtypes = [ci, cf, ci, ci, ci]
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extractor_tables.argtypes = [vp] + [vp] * 6
        L.orc_level_size.argtypes = [vp, ci, ci, ci, vp, vp]
        L.orc_extract.restype = ci
        L.orc_extract.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci]
        for f in ("orc_get_level", "orc_get_blur", "orc_get_score"):
            getattr(L, f).restype = ci
            getattr(L, f).argtypes = [vp, ci, vp]
        for f in ("orc_get_candidates", "orc_get_distributed"):
            getattr(L, f).restype = ci
            getattr(L, f).argtypes = [vp, ci, vp, ci]
        L.orc_search_radius.restype = ci
        L.orc_resize_linear_u8.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci]
        L.orc_gaussian7_s2_u8.argtypes = [vp, ci, ci, ci, vp, ci]
        L.orc_fast_score_u8.argtypes = [vp, ci, ci, ci, vp]
        L.orc_fast_nms.restype = ci
        L.orc_fast_nms.argtypes = [vp, ci, ci, ci, ci, vp, ci]
        L.orc_fast_atan2.restype = cf
        L.orc_fast_atan2.argtypes = [cf, cf]
        L.orc_distribute.restype = ci
        L.orc_distribute.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, ci]
        L.orc_distribute_ex.restype = ci
        L.orc_distribute_ex.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, ci, vp]
        L.orc_get_tie_cut.restype = ci
        L.orc_get_tie_cut.argtypes = [vp, ci]
        L.orc_hamming256.restype = ci
        L.orc_hamming256.argtypes = [vp, vp]
        L.orc_stereo_match.argtypes = [vp, vp, ci, vp, vp, ci, cd, cd, cd, vp, vp]
        L.orc_projection_match.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, ci, cd, cd, vp, vp]
        L.orc_projection_match_grid.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, ci, cd, cd, vp, vp]
        L.orc_projection_match_se3.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, ci, cd, cd, vp, vp]
        L.orc_projection_match_grid_se3.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, ci, cd, cd, vp, vp]
        L.orc_se3_apply.argtypes = [vp, vp, ci, vp]
        L.orc_reprojection_error_se3.argtypes = [vp, vp, vp, ci, vp, vp, vp]
        L.orc_track_pair.argtypes = [vp, cd, vp, cd, cd, vp, vp, ci, vp, vp, vp, vp, ci, vp, vp, ci]
        L.orc_stereo_sequence.restype = C.c_int64
        L.orc_stereo_sequence.argtypes = [vp, vp, ci, ci, ci, ci, ci, cf, ci, ci, ci, vp, cd, cd, vp, vp]
        L.orc_reprojection_error.argtypes = [vp, vp, vp, ci, vp, vp, vp]
        L.orc_knn2.argtypes = [vp, ci, vp, C.c_int64, C.c_int64, vp]
        L.orc_knn2_mt.argtypes = [vp, ci, vp, C.c_int64, C.c_int64, vp, ci]
        L.orc_stereo_frames.restype = C.c_int64
        L.orc_stereo_frames.argtypes = [vp, vp, ci, ci, ci, ci, ci, cf, ci, ci, ci, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Extractor:
    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.h = self.L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("bad extractor parameters")
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        self.L.orc_extractor_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def level_size(self, w, h, level):
        lw, lh = C.c_int(), C.c_int()
        self.L.orc_level_size(self.h, w, h, level, C.byref(lw), C.byref(lh))
        return lw.value, lh.value

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape if img.ndim == 2 else (0, 0)
        cap = self.nfeatures + 64 * self.nlevels + 64  # a level keeps all 4*n_ini first-split nodes even over its quota
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract(self.h, _p(img) if img.size else None, w, h, w, _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError(f"orc_extract failed: {n}")
        self._wh = (w, h)
        return kps[:n].copy(), desc[:n].copy()

    def _plane(self, fn, level):
        lw, lh = self.level_size(*self._wh, level)
        out = np.zeros((lh, lw), np.uint8)
        if fn(self.h, level, _p(out)) != 0:
            return None
        return out

    def level(self, l):
        return self._plane(self.L.orc_get_level, l)

    def blur(self, l):
        return self._plane(self.L.orc_get_blur, l)

    def score(self, l):
        return self._plane(self.L.orc_get_score, l)

    def candidates(self, l, cap=1 << 16):
        out = np.zeros((cap, 3), np.float32)
        n = self.L.orc_get_candidates(self.h, l, _p(out), cap)
        return out[:n].copy()

    def tie_cut(self, l):
        """True when level l's quadtree stopped between two equally-full nodes (heap-address order decides in the reference)."""
        return bool(self.L.orc_get_tie_cut(self.h, l))

    def distributed(self, l, cap=1 << 16):
        out = np.zeros((cap, 3), np.float32)
        n = self.L.orc_get_distributed(self.h, l, _p(out), cap)
        return out[:n].copy()


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orc_gaussian7_s2_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1])
    return dst


def fast_score(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().orc_fast_score_u8(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out))
    return out


def fast_nms(img, threshold, cap=1 << 16):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((cap, 3), np.float32)
    n = lib().orc_fast_nms(_p(img), img.shape[1], img.shape[0], img.shape[1], threshold, _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return np.float32(lib().orc_fast_atan2(float(y), float(x)))


def distribute(xyr, min_x, max_x, min_y, max_y, n_want, with_tie_cut=False):
    xyr = np.ascontiguousarray(xyr, np.float32)
    out = np.zeros((max(len(xyr), 1), 3), np.float32)
    cut = C.c_int(0)
    n = lib().orc_distribute_ex(_p(xyr), len(xyr), min_x, max_x, min_y, max_y, n_want, _p(out), len(out), C.byref(cut))
    if n < 0:
        raise ValueError("distribute failed")
    return (out[:n].copy(), bool(cut.value)) if with_tie_cut else out[:n].copy()


def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming256(_p(a), _p(b))


def stereo_match(kl, dl, kr, dr, y_thr=3.0, max_dx=100.0, ratio=0.5):
    kl, kr = np.ascontiguousarray(kl), np.ascontiguousarray(kr)
    dl, dr = np.ascontiguousarray(dl, np.uint8), np.ascontiguousarray(dr, np.uint8)
    idx = np.full(len(kl), -1, np.int32)
    dist = np.full(len(kl), -1, np.int32)
    lib().orc_stereo_match(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), y_thr, max_dx, ratio,
                           _p(idx), _p(dist))
    return idx, dist


def make_camera(fx, fy, cx, cy, d, width, height):
    cam = Camera()
    cam.fx, cam.fy, cam.cx, cam.cy = fx, fy, cx, cy
    for i in range(4):
        cam.d[i] = d[i]
    cam.width, cam.height = width, height
    return cam


def se3_apply(qt, x):
    """Xc = t + q * x the way g2o::SE3Quat / Eigen evaluate it; qt = (qx, qy, qz, qw, tx, ty, tz)."""
    qt = np.ascontiguousarray(qt, np.float64)
    x = np.ascontiguousarray(x, np.float64).reshape(-1, 3)
    out = np.zeros_like(x)
    lib().orc_se3_apply(_p(qt), _p(x), len(x), _p(out))
    return out


def projection_match(xw, mp_desc, skip, rt, cam, kps, kp_desc, radius, ratio=0.5, grid=False):
    """rt: 12 values = row-major 3x4 [R|t]; 7 values = g2o::SE3Quat (qx, qy, qz, qw, tx, ty, tz)."""
    xw = np.ascontiguousarray(xw, np.float64)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    skip = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    rt = np.ascontiguousarray(rt, np.float64).reshape(-1)
    assert rt.size in (7, 12)
    kps = np.ascontiguousarray(kps)
    kp_desc = np.ascontiguousarray(kp_desc, np.uint8)
    m = len(kps)
    to_q = np.full(m, -1, np.int32)
    dist = np.full(m, -1, np.int32)
    if rt.size == 7:
        fn = lib().orc_projection_match_grid_se3 if grid else lib().orc_projection_match_se3
    else:
        fn = lib().orc_projection_match_grid if grid else lib().orc_projection_match
    fn(_p(xw), _p(mp_desc), _p(skip), len(xw), _p(rt), C.byref(cam), _p(kps), _p(kp_desc), m, radius, ratio, _p(to_q), _p(dist))
    return to_q, dist


def track_pair(cam, baseline, rt, radius, kl_prev, dl_prev, kr_prev, sidx_prev, kps, desc, ratio=0.5, grid=False):
    """Tracking step between consecutive stereo frames (orc_track_pair) -> (track_idx, track_dist) per keypoint of `kps`."""
    kl_prev, kr_prev, kps = np.ascontiguousarray(kl_prev), np.ascontiguousarray(kr_prev), np.ascontiguousarray(kps)
    dl_prev, desc = np.ascontiguousarray(dl_prev, np.uint8), np.ascontiguousarray(desc, np.uint8)
    sidx_prev = np.ascontiguousarray(sidx_prev, np.int32)
    rt = np.ascontiguousarray(np.asarray(rt, np.float64)[:3, :4]).reshape(12)
    tidx = np.full(len(kps), -1, np.int32)
    tdist = np.full(len(kps), -1, np.int32)
    lib().orc_track_pair(C.byref(cam), baseline, _p(rt), radius, ratio, _p(kl_prev), _p(dl_prev), len(kl_prev), _p(kr_prev),
                         _p(sidx_prev), _p(kps), _p(desc), len(kps), _p(tidx), _p(tdist), int(grid))
    return tidx, tdist


def normalized_undistort(cam, kps):
    kps = np.ascontiguousarray(kps)
    out = np.zeros((len(kps), 2), np.float64)
    lib().orc_normalized_undistort(C.byref(cam), _p(kps), len(kps), _p(out))
    return out


def stereo_depth(cam, baseline, kps_l, norm_xy, kps_r, stereo_idx):
    kps_l, kps_r = np.ascontiguousarray(kps_l), np.ascontiguousarray(kps_r)
    norm_xy = np.ascontiguousarray(norm_xy, np.float64)
    stereo_idx = np.ascontiguousarray(stereo_idx, np.int32)
    xc = np.zeros((len(kps_l), 3), np.float64)
    valid = np.zeros(len(kps_l), np.uint8)
    lib().orc_stereo_depth(C.byref(cam), C.c_double(baseline), _p(kps_l), _p(norm_xy), len(kps_l), _p(kps_r), _p(stereo_idx),
                           _p(xc), _p(valid))
    return xc, valid


def reprojection_error(cam, rt, kps, xw, has_mp):
    kps = np.ascontiguousarray(kps)
    xw = np.ascontiguousarray(xw, np.float64)
    has_mp = np.ascontiguousarray(has_mp, np.uint8)
    err = np.zeros(len(kps), np.float64)
    if np.size(rt) == 7:
        qt = np.ascontiguousarray(rt, np.float64).reshape(7)
        lib().orc_reprojection_error_se3(C.byref(cam), _p(qt), _p(kps), len(kps), _p(xw), _p(has_mp), _p(err))
        return err
    rt = np.ascontiguousarray(np.asarray(rt, np.float64)[:3, :4]).reshape(12)
    lib().orc_reprojection_error(C.byref(cam), _p(rt), _p(kps), len(kps), _p(xw), _p(has_mp), _p(err))
    return err


def search_radius(kps, u, v, radius, cap=4096):
    kps = np.ascontiguousarray(kps)
    idx = np.zeros(cap, np.int32)
    n = lib().orc_search_radius(_p(kps), len(kps), C.c_double(u), C.c_double(v), C.c_double(radius), _p(idx), cap)
    return idx[:min(n, cap)].copy(), n


def search_nearest(kps, u, v):
    kps = np.ascontiguousarray(kps)
    i, d = C.c_int(), C.c_double()
    lib().orc_search_nearest(_p(kps), len(kps), C.c_double(u), C.c_double(v), C.byref(i), C.byref(d))
    return i.value, d.value


def vocab_transform(parent, is_leaf, node_desc, node_weight, L, features, levelsup):
    parent = np.ascontiguousarray(parent, np.int32)
    is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
    node_desc = np.ascontiguousarray(node_desc, np.uint8)
    node_weight = np.ascontiguousarray(node_weight, np.float64)
    features = np.ascontiguousarray(features, np.uint8)
    n = len(features)
    wid, w, nid = np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(n, np.int32)
    lib().orc_vocab_transform(len(parent), _p(parent), _p(is_leaf), _p(node_desc), _p(node_weight), L, _p(features), n, levelsup,
                              _p(wid), _p(w), _p(nid))
    return wid, w, nid


def knn2(queries, db, idx_base=0, nthreads=1):
    queries = np.ascontiguousarray(queries, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    out = np.zeros((len(queries), 4), np.int32)
    lib().orc_knn2_mt(_p(queries), len(queries), _p(db), len(db), idx_base, _p(out), int(nthreads))
    return out


def stereo_frames(left, right, nthreads, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    count, h, w = left.shape
    tot = C.c_int64()
    m = lib().orc_stereo_frames(_p(left), _p(right), count, w, h, nthreads, nfeatures, scale_factor,
                                nlevels, ini_th, min_th, C.byref(tot))
    return int(m), int(tot.value)


def stereo_sequence(left, right, nthreads, cam, baseline, radius=50.0, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20,
                    min_th=7):
    """CPU baseline of the sequence workload -> (stereo matches, keypoints, tracked keypoints)."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    count, h, w = left.shape
    tot, trk = C.c_int64(), C.c_int64()
    m = lib().orc_stereo_sequence(_p(left), _p(right), count, w, h, nthreads, nfeatures, scale_factor, nlevels, ini_th, min_th,
                                  C.byref(cam), baseline, radius, C.byref(tot), C.byref(trk))
    return int(m), int(tot.value), int(trk.value)
