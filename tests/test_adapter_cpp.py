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


def _build():
    import __graft_entry__ as g
    if not os.path.exists(api.LIB_PATH):
        g.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)


def _run(tmp_path, L, R):
    L.tofile(tmp_path / "l.raw")
    R.tofile(tmp_path / "r.raw")
    p = subprocess.run([EXE, str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), str(L.shape[1]), str(L.shape[0])],
                       capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout.strip()


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
    rc, out = _run(tmp_path, L, R)
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
    # every triangulated point re-projects onto its own keypoint; most are accepted by the ratio test
    assert int(got["proj"]) > 0.5 * (si >= 0).sum() and int(got["self"]) > 0.9 * int(got["proj"])
