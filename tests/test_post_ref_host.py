"""CPU suite: oracle/post_oracle.py (the numpy checker of csrc/dvr_post.cu) against the reference's own
frame post passes — tsd/src/render_pipeline/passes/{OutlineRenderPass,VisualizeDepthPass,AnariSceneRenderPass}.cpp
compiled in place into oracle/_ref/libref_post.so (serial parallel_for fallback; recipe oracle/Makefile, wrappers
oracle/ref_post/).  Byte / integer work: bit-exact.  With this the GPU post-pass tests (tests/test_gpu_post.py, CUDA vs
post_oracle.py) are pinned to reference code; what stays restated is named in oracle/ref_post/shim/tsd/core/TSDMath.hpp
(the two helium colour-conversion helpers and linalg's lerp: the ANARI-SDK is not in this image)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import post_oracle as PO  # noqa: E402

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_post.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_post.so not built")

_lib = None


def ref():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        u32p, f32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        _lib.refpost_outline.argtypes = [u32p, u32p, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.refpost_visualize_depth.argtypes = [u32p, f32p, C.c_uint32, C.c_uint32, C.c_float]
        _lib.refpost_convert_float_color.argtypes = [f32p, u8p, C.c_size_t]
        _lib.refpost_composite.argtypes = [u32p, f32p, u32p, u32p, f32p, u32p, C.c_uint32, C.c_uint32, C.c_int]
        for f in (_lib.refpost_outline, _lib.refpost_visualize_depth, _lib.refpost_convert_float_color, _lib.refpost_composite):
            f.restype = None
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


@pytest.mark.parametrize("n", [1, 255, 640 * 480])
def test_convert_float_color_equals_the_reference_pass(n):
    rng = np.random.default_rng(n)
    src = (rng.random((n, 4)) * 1.4 - 0.2).astype(np.float32)
    src[0] = (0.0, 1.0, 0.999999, 2.0)
    if n > 4:
        src[1] = (-0.0, 1.0 / 255.0, 254.999 / 255.0, 0.5)
        src[2] = (np.float32(1.0) - np.float32(2.0 ** -24), 0.00392, 0.00393, -5.0)
    out = np.zeros(n * 4, np.uint8)
    ref().refpost_convert_float_color(_p(src, C.c_float), _p(out, C.c_uint8), n * 4)
    assert np.array_equal(out.view(np.uint32), PO.convert_float_color(src))


@pytest.mark.parametrize("first,with_ids", [(True, True), (False, True), (False, False)])
def test_composite_depth_equals_the_reference_pass(first, with_ids):
    rng = np.random.default_rng(5)
    w, h = 211, 97
    n = w * h
    co, ci = rng.integers(0, 2 ** 32, n, dtype=np.uint32), rng.integers(0, 2 ** 32, n, dtype=np.uint32)
    do, di = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    do[::7] = np.inf
    di[::11] = np.inf
    di[5], do[6] = np.nan, np.nan  # comparisons with NaN are false: the pixel is kept
    di[8] = do[8]                  # equal depth: kept (strict <)
    io, ii = rng.integers(0, 50, n, dtype=np.uint32), rng.integers(0, 50, n, dtype=np.uint32)
    wc, wd, wi = PO.composite_depth(co, do, io, ci, di, ii if with_ids else None, first)
    rc, rd, ri = co.copy(), do.copy(), io.copy()
    ref().refpost_composite(_p(rc, C.c_uint32), _p(rd, C.c_float), _p(ri, C.c_uint32), _p(ci, C.c_uint32), _p(di, C.c_float),
                            _p(ii, C.c_uint32) if with_ids else None, w, h, int(first))
    assert np.array_equal(rc, wc) and np.array_equal(rd, wd, equal_nan=True) and np.array_equal(ri, wi)


@pytest.mark.parametrize("w,h", [(64, 48), (1, 1), (3, 2), (2, 7), (257, 131)])
def test_outline_equals_the_reference_pass(w, h):
    """incl. the reference's unsigned `max(0u, y - 1)`: row 0 and column 0 are never outlined"""
    rng = np.random.default_rng(w * h)
    ids = np.full((h, w), 0xFFFFFFFF, np.uint32)
    yy, xx = np.mgrid[0:h, 0:w]
    ids[(xx - w * 0.4) ** 2 + (yy - h * 0.5) ** 2 < (min(w, h) * 0.3) ** 2] = 7
    ids[: max(h // 5, 1), : max(w // 3, 1)] = 7  # touches row 0 / column 0
    ids[rng.random((h, w)) < 0.02] = 3
    color = rng.integers(0, 2 ** 32, w * h, dtype=np.uint32)
    for oid in (7, 3, 0xFFFFFFFF):
        want = PO.outline(color, ids.ravel(), w, h, oid)
        got = color.copy()
        ref().refpost_outline(_p(got, C.c_uint32), _p(np.ascontiguousarray(ids.ravel()), C.c_uint32), w, h, oid)
        if oid == 0xFFFFFFFF:  # OutlineRenderPass::render: ~0u means "no outline"
            assert np.array_equal(got, color)
        else:
            assert np.array_equal(got, want)


def test_shade_pixel_equals_the_reference_pass_on_every_byte_value():
    # a 3x3 image whose centre has count 4: the centre pixel goes through shadePixel; sweep all byte values per channel
    ids = np.zeros((3, 3), np.uint32)
    ids[1:3, 1:3] = 1
    for v in range(256):
        c = np.uint32(v | ((255 - v) << 8) | (((v * 7) & 255) << 16) | (((v * 13 + 5) & 255) << 24))
        color = np.full(9, c, np.uint32)
        got = color.copy()
        ref().refpost_outline(_p(got, C.c_uint32), _p(np.ascontiguousarray(ids.ravel()), C.c_uint32), 3, 3, 1)
        assert got[4] == PO.shade_pixel(np.array([c], np.uint32))[0]
        assert np.array_equal(got, PO.outline(color, ids.ravel(), 3, 3, 1))


@pytest.mark.parametrize("max_depth", [1.0, 6.0, 1e30, 0.37])
def test_visualize_depth_equals_the_reference_pass(max_depth):
    rng = np.random.default_rng(11)
    w, h = 97, 33
    depth = (rng.random(w * h) * max_depth * 1.3).astype(np.float32)
    depth[:6] = (0.0, np.inf, max_depth, -1.0, np.float32(max_depth) * np.float32(0.5), 3.4028235e38)
    got = np.zeros(w * h, np.uint32)
    ref().refpost_visualize_depth(_p(got, C.c_uint32), _p(depth, C.c_float), w, h, max_depth)
    assert np.array_equal(got, PO.visualize_depth(depth, max_depth))
