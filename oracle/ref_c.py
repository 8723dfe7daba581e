"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libslamref.so: the REFERENCE'S OWN
src/orb_extractor.cpp, src/matcher.cpp and src/camera.cpp, compiled unmodified from /root/reference by
oracle/ref_build/Makefile against stand-in third-party headers.

Used by tests/ to pin oracle/orb_oracle.c (and through it the CUDA path) to the reference's code, by
tools/gen_golden.py to generate tests/golden/, and by bench.py --impl reference.  Never by the product package.
The library is built in the authoring container (where /root/reference exists) and travels to the GPU box as a
prebuilt file; `available()` is False when neither the library nor the reference sources are there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

try:  # same 28-byte keypoint / camera structs as the C oracle's binding
    from .oracle_c import KP_DTYPE, Camera, make_camera  # noqa: F401
except ImportError:  # tests put oracle/ itself on sys.path
    from oracle_c import KP_DTYPE, Camera, make_camera  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libslamref.so")
REF_ROOT = os.environ.get("SLAM_REFERENCE_ROOT", "/root/reference")
HEAP_MALLOC, HEAP_MONOTONIC = 0, 1


def build(force: bool = False) -> str | None:
    """Build oracle/_ref when the reference sources are present; otherwise use the prebuilt library if any."""
    if os.path.isdir(os.path.join(REF_ROOT, "src")):
        cmd = ["make", "-C", os.path.join(_HERE, "ref_build"), f"REF={REF_ROOT}"] + (["-B"] if force else ["-s"])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return _LIB if os.path.exists(_LIB) else None


def available() -> bool:
    try:
        return build() is not None
    except (subprocess.CalledProcessError, OSError):
        return os.path.exists(_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libslamref.so is not built and /root/reference is absent")
        L = C.CDLL(_LIB)
        vp, ci, cf, cd = C.c_void_p, C.c_int, C.c_float, C.c_double
        L.ref_extractor_create.restype = vp
        L.ref_extractor_create.argtypes = [ci, cf, ci, ci, ci]
        L.ref_extractor_destroy.argtypes = [vp]
        L.ref_extractor_tables.argtypes = [vp] * 7
        L.ref_extract.restype = ci
        L.ref_extract.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci]
        L.ref_level_size.restype = ci
        L.ref_level_size.argtypes = [vp, ci, vp, vp]
        L.ref_get_level.restype = ci
        L.ref_get_level.argtypes = [vp, ci, ci, vp]
        L.ref_get_blur.restype = ci
        L.ref_get_blur.argtypes = [vp, ci, vp]
        L.ref_get_candidates.restype = ci
        L.ref_get_candidates.argtypes = [vp, ci, vp, ci]
        L.ref_distribute.restype = ci
        L.ref_distribute.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, ci]
        L.ref_hamming256.restype = ci
        L.ref_hamming256.argtypes = [vp, vp]
        L.ref_fast_atan2.restype = cf
        L.ref_fast_atan2.argtypes = [cf, cf]
        L.ref_camera_project.restype = ci
        L.ref_camera_project.argtypes = [vp, vp, vp]
        L.ref_normalized_undistort.argtypes = [vp, vp, ci, vp]
        L.ref_se3_apply.argtypes = [vp, vp, ci, vp, vp]
        L.ref_stereo_match.argtypes = [vp, vp, ci, vp, vp, ci, vp, vp]
        L.ref_projection_match.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, ci, cd, vp]
        L.ref_stereo_sequence.restype = C.c_int64
        L.ref_stereo_sequence.argtypes = [vp, vp, ci, ci, ci, ci, ci, cf, ci, ci, ci, vp, cd, cd, vp, vp]
        L.ref_set_heap_mode.argtypes = [ci]
        L.ref_arena_allocations.restype = C.c_long
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_heap_mode(mode: int) -> None:
    """HEAP_MONOTONIC: quadtree list nodes get increasing addresses (= creation order, the oracle's rule T1);
    HEAP_MALLOC: glibc malloc decides, as in the reference binary."""
    lib().ref_set_heap_mode(mode)


class Extractor:
    """ORB_SLAM2::ORBextractor of the reference (include/orb_extractor.h:45-133)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.h = self.L.ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        self.L.ref_extractor_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def extract(self, img, cap=None):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = cap or (w * h // 16 + 4096)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.ref_extract(self.h, _p(img), w, h, w, _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError("ref_extract: capacity")
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        self.L.ref_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, l, ring=False):
        w, h = self.level_size(l)
        e = 19 if ring else 0
        out = np.zeros((h + 2 * e, w + 2 * e), np.uint8)
        return out if self.L.ref_get_level(self.h, l, int(ring), _p(out)) == 0 else None

    def blur(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h, w), np.uint8)
        return out if self.L.ref_get_blur(self.h, l, _p(out)) == 0 else None

    def candidates(self, l, cap=1 << 20):
        out = np.zeros((cap, 3), np.float32)
        n = self.L.ref_get_candidates(self.h, l, _p(out), cap)
        return out[:n].copy()

    def distribute(self, xyr, min_x, max_x, min_y, max_y, n_want, level=0):
        xyr = np.ascontiguousarray(xyr, np.float32)
        out = np.zeros((max(len(xyr), 1) + 8, 3), np.float32)
        n = self.L.ref_distribute(self.h, _p(xyr), len(xyr), min_x, max_x, min_y, max_y, n_want, level, _p(out), len(out))
        return out[:n].copy()


def hamming256(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return lib().ref_hamming256(_p(a), _p(b))


def fast_atan2(y, x):
    return np.float32(lib().ref_fast_atan2(float(y), float(x)))


def camera_project(cam, xc):
    xc = np.ascontiguousarray(xc, np.float64)
    uv = np.zeros(2, np.float64)
    inside = lib().ref_camera_project(C.byref(cam), _p(xc), _p(uv))
    return uv, bool(inside)


def normalized_undistort(cam, kps):
    kps = np.ascontiguousarray(kps)
    out = np.zeros((len(kps), 2), np.float64)
    lib().ref_normalized_undistort(C.byref(cam), _p(kps), len(kps), _p(out))
    return out


def se3_apply(qt, x):
    """g2o::SE3Quat(q, t) * x for qt = (qx, qy, qz, qw, tx, ty, tz) -> (points, quaternion the SE3Quat holds)."""
    qt = np.ascontiguousarray(qt, np.float64)
    x = np.ascontiguousarray(x, np.float64).reshape(-1, 3)
    out = np.zeros_like(x)
    q = np.zeros(4, np.float64)
    lib().ref_se3_apply(_p(qt), _p(x), len(x), _p(out), _p(q))
    return out, q


def stereo_match(kl, dl, kr, dr, cam):
    """StereoMatch(StereoFrame*) of src/matcher.cpp:54-132 -> right index or -1 per left keypoint."""
    kl, kr = np.ascontiguousarray(kl), np.ascontiguousarray(kr)
    dl, dr = np.ascontiguousarray(dl, np.uint8), np.ascontiguousarray(dr, np.uint8)
    idx = np.full(len(kl), -1, np.int32)
    lib().ref_stereo_match(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), C.byref(cam), _p(idx))
    return idx


def projection_match(xw, mp_desc, skip, qt, cam, kps, kp_desc, radius):
    """ProjectionMatch of src/matcher.cpp:134-209 -> map-point index or -1 per frame keypoint."""
    xw = np.ascontiguousarray(xw, np.float64)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    skip = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    qt = np.ascontiguousarray(qt, np.float64)
    kps = np.ascontiguousarray(kps)
    kp_desc = np.ascontiguousarray(kp_desc, np.uint8)
    to_q = np.full(len(kps), -1, np.int32)
    lib().ref_projection_match(_p(xw), _p(mp_desc), _p(skip), len(xw), _p(qt), C.byref(cam), _p(kps), _p(kp_desc), len(kps),
                               C.c_double(radius), _p(to_q))
    return to_q


def stereo_sequence(left, right, nthreads, cam, baseline, radius=50.0, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20,
                    min_th=7):
    """CPU baseline of the sequence workload on the reference's own code (ref_stereo_sequence) ->
    (stereo matches, keypoints, tracked keypoints).  glibc heap, as the reference runs."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    count, h, w = left.shape
    tot, trk = C.c_int64(), C.c_int64()
    m = lib().ref_stereo_sequence(_p(left), _p(right), count, w, h, nthreads, nfeatures, scale_factor, nlevels, ini_th, min_th,
                                  C.byref(cam), baseline, radius, C.byref(tot), C.byref(trk))
    return int(m), int(tot.value), int(trk.value)
