"""CPU: the host layer's PNG reader (VO::read_img reads image_0/%06d.png like the reference's cv::imread GRAYSCALE,
visual_odometry.cpp:42-51) against cv2.imread on files written by cv2 and on hand-built files using every PNG filter."""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stereo-visual-slam_b200", "libvslam_b200_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB):
        pytest.skip("host library not built")
    return C.CDLL(LIB)


def _read(host, path, cap=4 << 20):
    out = np.zeros(cap, np.uint8)
    w, h = C.c_int(0), C.c_int(0)
    st = host.vslam_host_read_png(str(path).encode(), out.ctypes.data_as(C.c_void_p), cap, C.byref(w), C.byref(h))
    return st, (out[:w.value * h.value].reshape(h.value, w.value) if st == 0 else None)


def _chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)


def _write_png(path, img, filters, ctype=0):
    h, w = img.shape[:2]
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    rows = img.reshape(h, w * ch).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        ft = filters[y % len(filters)]
        cur = rows[y]
        a = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        c = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
        if ft == 0: f = cur
        elif ft == 1: f = cur - a
        elif ft == 2: f = cur - prev
        elif ft == 3: f = cur - ((a + prev) >> 1)
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
            f = cur - pred
        raw += bytes([ft]) + (f & 255).astype(np.uint8).tobytes()
        prev = cur
    data = zlib.compress(bytes(raw), 6)
    with open(path, "wb") as fo:
        fo.write(b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
                 _chunk(b"IDAT", data[:len(data) // 2]) + _chunk(b"IDAT", data[len(data) // 2:]) + _chunk(b"IEND", b""))


def test_png_written_by_cv2(host, pkg, tmp_path):
    left, _, _ = pkg.synth.synth_pair(2)
    p = tmp_path / "000000.png"
    cv2.imwrite(str(p), left)
    st, img = _read(host, p)
    assert st == 0 and np.array_equal(img, left) and np.array_equal(img, cv2.imread(str(p), cv2.IMREAD_GRAYSCALE))


def test_every_filter_type_and_split_idat(host, tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    for filters in ([0], [1], [2], [3], [4], [0, 1, 2, 3, 4]):
        p = tmp_path / f"f{''.join(map(str, filters))}.png"
        _write_png(p, img, filters)
        st, out = _read(host, p)
        assert st == 0 and np.array_equal(out, img) and np.array_equal(out, cv2.imread(str(p), cv2.IMREAD_GRAYSCALE))


def test_colour_png_is_reduced_like_cv2(host, tmp_path):
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (120, 131, 3), dtype=np.uint8)
    p = tmp_path / "rgb.png"
    _write_png(p, rgb, [4, 1], ctype=2)
    st, out = _read(host, p)
    assert st == 0 and np.array_equal(out, cv2.imread(str(p), cv2.IMREAD_GRAYSCALE))


def test_rejects_what_it_does_not_support(host, tmp_path):
    p = tmp_path / "bad.png"
    p.write_bytes(b"not a png")
    assert _read(host, p)[0] == -1
    img16 = np.zeros((4, 4), np.uint16)
    p16 = tmp_path / "d16.png"
    cv2.imwrite(str(p16), img16)
    assert _read(host, p16)[0] == -1
    assert _read(host, tmp_path / "missing.png")[0] == -1
