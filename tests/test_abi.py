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


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (Python, CUDA, C++ host layer) may import, include or
    execute it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may."""
    pk = os.path.join(ROOT, "stereo-visual-slam_b200")
    offenders = []
    for dirpath, _, files in os.walk(pk):
        if os.sep + "build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".inc")):
                continue
            txt = open(os.path.join(dirpath, f), errors="ignore").read()
            for m in re.finditer(r"^\s*(from\s+oracle|import\s+oracle|#include\s+[\"<][^\">]*oracle)", txt, flags=re.M):
                offenders.append((os.path.relpath(os.path.join(dirpath, f), ROOT), m.group(0).strip()))
            if re.search(r"oracle/_build|oracle/_ref|libba_oracle", txt):
                offenders.append((os.path.relpath(os.path.join(dirpath, f), ROOT), "links the oracle library"))
    assert not offenders, offenders


def test_host_pool_without_a_device_and_foreign_pointers(pkg):
    """vslam_host_alloc hands out pinned blocks or NULL (no device: this container), never throws; vslam_host_free
    ignores NULL and pointers the pool did not hand out; a freed block is handed out again for the same size."""
    import ctypes as C
    lib = pkg.ffi.load_library()
    lib.vslam_host_free(None)
    foreign = C.create_string_buffer(64)
    lib.vslam_host_free(C.cast(foreign, C.c_void_p))  # not ours: ignored
    p = lib.vslam_host_alloc(1 << 20)
    if p:  # a CUDA device is present
        C.memset(p, 0x5A, 1 << 20)
        lib.vslam_host_free(p)
        q = lib.vslam_host_alloc((1 << 20) - 100)  # same 64 KiB size class: the block comes back
        assert q == p
        lib.vslam_host_free(q)
