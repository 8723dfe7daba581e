"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/sfe.h
declares, the Python binding covers exactly that set, and compute entry points fail loudly
(no CPU fallback) when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as graft
from slam_toolkit_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(api.LIB_PATH):
        graft.build()
    return api.lib()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "sfe.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(sfe_[a-z0-9_]+)\s*\(", hdr))


def test_every_declared_symbol_is_exported_and_bound(built):
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(built, n), f"libsfe.so does not export {n}"
    assert names == set(api.SIGNATURES), names ^ set(api.SIGNATURES)


def test_keypoint_layout_is_cv_keypoint():
    assert api.KP_DTYPE.itemsize == 28
    assert [api.KP_DTYPE.fields[k][1] for k in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_host_only_entry_points(built, oracle):
    assert built.sfe_abi_version() == 2
    assert built.sfe_status_string(0) == b"ok"
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert api.hamming256(a, b) == oracle.hamming256(a, b) == int(np.unpackbits(a ^ b).sum())


def test_bad_arguments_are_rejected(built):
    h = C.c_void_p()
    bad = api.ExtractorParams(2000, 1.2, 99, 20, 7)
    assert built.sfe_extractor_create(C.byref(bad), 0, 2, C.byref(h)) == api.SFE_ERR_BAD_ARG
    assert built.sfe_extractor_create(None, 0, 2, C.byref(h)) == api.SFE_ERR_BAD_ARG
    assert b"nlevels" in built.sfe_last_error() or b"null" in built.sfe_last_error()
    assert built.sfe_matcher_create(0, None) == api.SFE_ERR_BAD_ARG


def test_no_cpu_fallback(built):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.SfeError) as e:
        api.ORBextractor()
    assert e.value.status in (api.SFE_ERR_NO_DEVICE, api.SFE_ERR_CUDA)
    with pytest.raises(api.SfeError):
        api.Matcher()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "slam-toolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                for line in open(os.path.join(dirpath, f)):
                    code = line.split("//")[0].split("#", 1)[0] if not line.lstrip().startswith("#include") else line
                    uses = re.search(r"#include.*oracle|import\s+oracle|from\s+oracle|liborb_oracle|orc_[a-z_]+\s*\(", code)
                    assert not uses, f"{f}: {line.strip()}"
    inc = open(os.path.join(ROOT, "include", "sfe.h")).read()
    assert "orb_oracle" not in inc


def test_pattern_tables_identical():
    a = open(os.path.join(ROOT, "oracle", "orb_pattern_data.inc")).read()
    b = open(os.path.join(ROOT, "slam-toolkit_b200", "csrc", "orb_pattern_data.inc")).read()
    assert a == b
