"""include/sfe_adapter.hpp compiled against stand-in OpenCV types and mock Frame / Mappoint classes
that expose the reference's accessor names (tests/cpp/adapter_test.cpp).  CPU: it must compile, link
and fail loudly without a device.  GPU: its results must equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

from slam_toolkit_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_test")


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _fnv_more(h: int, b: bytes) -> int:
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _build():
    import __graft_entry__ as g
    if not os.path.exists(api.LIB_PATH):
        g.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)


def _run(tmp_path, L, R, voc=None):
    L.tofile(tmp_path / "l.raw")
    R.tofile(tmp_path / "r.raw")
    p = subprocess.run([EXE, str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), str(L.shape[1]), str(L.shape[0])] +
                       ([str(voc)] if voc else []), capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout.strip()


def _write_vocabulary(path, k, voc):
    """DBoW2 text format (TemplatedVocabulary::saveToTextFile): 'k L scoring weighting', then per node
    'parent isLeaf d0 .. d31 weight'."""
    parent, is_leaf, desc, weight, L = voc
    with open(path, "w") as f:
        f.write(f"{k} {L} 0 0\n")
        for i in range(1, len(parent)):
            f.write(f"{parent[i]} {is_leaf[i]} " + " ".join(str(int(b)) for b in desc[i]) + f" {float(weight[i])!r}\n")


def test_adapter_compiles_and_has_no_cpu_fallback(tmp_path):
    _build()
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    rc, out = _run(tmp_path, *synth.stereo_pair(0, 160, 120))
    assert rc == 1 and "exception" in out and "sfe_extractor_create" in out


@pytest.mark.gpu
def test_adapter_results_equal_oracle(tmp_path, oracle):
    _build()
    L, R = synth.stereo_pair(2)
    voc = synth.vocabulary(k=10, L=3, seed=5)
    _write_vocabulary(tmp_path / "voc.txt", 10, voc)
    rc, out = _run(tmp_path, L, R, tmp_path / "voc.txt")
    assert rc == 0, out
    got = dict(kv.split("=") for kv in out.split())
    ref = oracle.Extractor()
    kl, dl = ref.extract(L)
    kr, dr = ref.extract(R)
    si, _ = oracle.stereo_match(kl, dl, kr, dr)
    assert (int(got["nl"]), int(got["nr"])) == (len(kl), len(kr))
    assert int(got["kps"], 16) == _fnv(kl.tobytes()) and int(got["desc"], 16) == _fnv(dl.tobytes())
    assert int(got["stereo"], 16) == _fnv(si.astype(np.int32).tobytes())
    assert int(got["dd"]) == oracle.hamming256(dl[0], dl[1])
    # Frame::ComputeBoW through sfe_adapter::Vocabulary (text loader + transform) == oracle descent + literal assembly
    from test_oracle_matchers import _bow_literal
    wid, w, nid = oracle.vocab_transform(voc[0], voc[1], voc[2], voc[3], voc[4], dl, 4)
    lit = _bow_literal(wid, w, 0, 1)
    hb = 1469598103934665603
    for k_, v_ in lit:
        hb = _fnv_more(hb, np.uint32(k_).tobytes() + np.float64(v_).tobytes())
    fv = {}
    for i in np.nonzero(w > 0)[0]:
        fv.setdefault(int(nid[i]), []).append(int(i))
    hf = 1469598103934665603
    for k_ in sorted(fv):
        hf = _fnv_more(hf, np.uint32(k_).tobytes() + np.array(fv[k_], np.uint32).tobytes())
    assert int(got["bown"]) == len(lit) and int(got["bow"], 16) == hb and int(got["fv"], 16) == hf
    # every triangulated point re-projects onto its own keypoint; most are accepted by the ratio test
    assert int(got["proj"]) > 0.5 * (si >= 0).sum() and int(got["self"]) > 0.9 * int(got["proj"])
    # sfe_comm_create_local + sfe_knn2_sharded + sfe_projection_match_sharded called from C++ on every visible GPU
    assert int(got["sharded_gpus"]) >= 1 and got["sharded_knn"] == "ok" and got["sharded_proj"] == "ok"


@pytest.mark.gpu
def test_adapter_fused_stereo_call_and_graph_replay(tmp_path):
    """ORBextractor::extractStereo (one round trip) returns the bytes of extract + extract + StereoMatch; both run their
    kernels as replayed CUDA graphs from the third call on (tests/cpp/adapter_latency.cpp compares them after 320 calls)."""
    _build()
    L, R = synth.stereo_pair(3)
    L.tofile(tmp_path / "l.raw")
    R.tofile(tmp_path / "r.raw")
    p = subprocess.run([os.path.join(ROOT, "tests", "cpp", "adapter_latency"), str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                        str(L.shape[1]), str(L.shape[0]), "50"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    got = dict(kv.split("=") for kv in p.stdout.split())
    assert got["same"] == "1" and int(got["nl"]) > 1900
