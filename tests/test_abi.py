"""CPU: the C-ABI library loads and exports every symbol include/vslam_b200.h declares (no compute)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vslam_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vslam_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(pkg):
    lib = pkg.ffi.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vslam_b200.h but not exported"


def test_binding_covers_header(pkg):
    assert sorted(pkg.ffi.SIGNATURES) == _declared_symbols()


def test_abi_version_and_status(pkg):
    lib = pkg.ffi.load_library()
    assert lib.vslam_abi_version() == 1
    assert lib.vslam_status_string(0) == b"ok"
    assert b"capacity" in lib.vslam_status_string(-2)


def test_pod_layouts(pkg):
    # cv::KeyPoint is 28 bytes, cv::DMatch 16
    assert pkg.KEYPOINT_DTYPE.itemsize == 28
    assert pkg.DMATCH_DTYPE.itemsize == 16


def test_no_cpu_fallback(pkg):
    """Without a GPU the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.VslamError) as e:
        pkg.Context()
    assert e.value.status == -4
