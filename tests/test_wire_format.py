"""SFER byte stream (include/sfe.h): pack / unpack round trips, the size contract and the rejection of
damaged streams.  Host-only code of libsfe.so: runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from slam_toolkit_b200 import api


def _fake_results(frames, cap, rng, stereo=True, track=True):
    out = {"kps_l": np.zeros((frames, cap), api.KP_DTYPE), "desc_l": rng.integers(0, 256, (frames, cap, 32), dtype=np.uint8),
           "n_l": rng.integers(0, cap + 1, frames).astype(np.int32)}
    for k in ("x", "y", "size", "angle", "response"):
        out["kps_l"][k] = rng.uniform(0, 1000, (frames, cap)).astype(np.float32)
    out["kps_l"]["octave"] = rng.integers(0, 8, (frames, cap))
    out["kps_l"]["class_id"] = -1
    if stereo:
        out["kps_r"] = np.roll(out["kps_l"], 1, axis=1).copy()
        out["desc_r"] = rng.integers(0, 256, (frames, cap, 32), dtype=np.uint8)
        out["n_r"] = rng.integers(0, cap + 1, frames).astype(np.int32)
        out["stereo_idx"] = rng.integers(-1, cap, (frames, cap)).astype(np.int32)
        out["stereo_dist"] = rng.integers(-1, 257, (frames, cap)).astype(np.int32)
    if track:
        out["track_idx"] = rng.integers(-1, cap, (frames, cap)).astype(np.int32)
        out["track_dist"] = rng.integers(-1, 257, (frames, cap)).astype(np.int32)
    return out


def _valid_equal(a, b):
    assert np.array_equal(a["n_l"], b["n_l"])
    for f in range(len(a["n_l"])):
        nl = a["n_l"][f]
        for k in ("kps_l", "desc_l", "stereo_idx", "stereo_dist", "track_idx", "track_dist"):
            if k in a:
                assert np.array_equal(a[k][f, :nl], b[k][f, :nl]), (k, f)
        if "n_r" in a:
            nr = a["n_r"][f]
            assert nr == b["n_r"][f]
            assert np.array_equal(a["kps_r"][f, :nr], b["kps_r"][f, :nr]) and np.array_equal(a["desc_r"][f, :nr], b["desc_r"][f, :nr])


@pytest.mark.parametrize("stereo,track", [(True, True), (True, False), (False, False), (False, True)])
def test_round_trip_and_size(stereo, track):
    rng = np.random.default_rng(3)
    out = _fake_results(5, 37, rng, stereo, track)
    out["n_l"][2] = 0                      # an empty frame
    data = api.pack_results(out, 1241, 376)
    nl, nr = out["n_l"].astype(np.int64), (out["n_r"] if stereo else np.zeros(5, np.int32)).astype(np.int64)
    expect = 48 + int((8 + nl * 60 + (nr * 60 + nl * 8 if stereo else 0) + (nl * 8 if track else 0)).sum())
    assert len(data) == expect
    back, wh = api.unpack_results(data)
    assert wh == (1241, 376) and set(back) == set(out)
    _valid_equal(out, back)
    assert back["kps_l"].shape[1] == max(int(nl.max()), int(nr.max()), 1)   # cap shrinks to the largest frame
    assert api.pack_results(back, 1241, 376) == data                      # canonical: re-packing reproduces the bytes


def test_zero_frames_and_damaged_streams():
    empty = {"kps_l": np.zeros((0, 4), api.KP_DTYPE), "desc_l": np.zeros((0, 4, 32), np.uint8), "n_l": np.zeros(0, np.int32)}
    data = api.pack_results(empty)
    assert len(data) == 48
    back, _ = api.unpack_results(data)
    assert back["kps_l"].shape[0] == 0
    rng = np.random.default_rng(4)
    good = api.pack_results(_fake_results(3, 20, rng))
    for bad in (good[:40], good[:-1], good + b"\0", b"XXXX" + good[4:], good[:100] + bytes([good[100] ^ 1]) + good[101:],
                good[:4] + (2).to_bytes(4, "little") + good[8:]):
        with pytest.raises(api.SfeError):
            api.unpack_results(bad)
    # pack into a buffer that is too small: capacity error, nothing silently truncated
    out = _fake_results(2, 8, rng)
    buf = np.empty(64, np.uint8)
    rc = api.lib().sfe_results_pack(buf.ctypes.data_as(C.c_void_p), buf.nbytes, 2, 8, 0, 0, 3,
                                    *[np.ascontiguousarray(out[k]).ctypes.data_as(C.c_void_p) for k in api._RES_KEYS], None)
    assert rc == api.SFE_ERR_CAPACITY
