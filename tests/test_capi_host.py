"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/dvr_b200.h declares, its
host-side helpers (camera set-up, TF discretisation) agree bit-for-bit with the oracle / the
reference's host code, and every compute entry fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_binding as ob
from conftest import HAS_GPU
from visrtx_b200 import capi, scenes
from test_oracle_golden import TF_CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_and_library_export_the_same_symbols():
    hdr = open(os.path.join(ROOT, "include", "dvr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dvr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(capi.lib, name), f"libdvr_b200.so does not export {name}"


def test_struct_layouts_match_header_sizes(tmp_path):
    """The ctypes mirrors against what the C compiler makes of include/dvr_b200.h itself (sizes and the offsets of the
    last members): a mismatch would corrupt every launch."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dvr_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu ",'
                   'sizeof(DvrCamera),sizeof(DvrVolumeInstance),sizeof(DvrFrameBuffers),sizeof(DvrFrameParams),'
                   'sizeof(DvrRenderStats),sizeof(DvrPeerSync),offsetof(DvrFrameParams,backgroundImage),'
                   'offsetof(DvrFrameParams,partialCullToBounds),offsetof(DvrFrameBuffers,depthMirror),sizeof(DvrSlabExchange),offsetof(DvrSlabExchange,timing));'
                   'printf("%zu %zu %zu %zu %zu\\n",sizeof(DvrSurfaceDesc),offsetof(DvrSurfaceDesc,objectToWorld),sizeof(DvrLight),sizeof(DvrSceneParams),offsetof(DvrSceneParams,cullTriangleBackfaces));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    want = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    got = [C.sizeof(capi.DvrCamera), C.sizeof(capi.DvrVolumeInstance), C.sizeof(capi.DvrFrameBuffers),
           C.sizeof(capi.DvrFrameParams), C.sizeof(capi.DvrRenderStats), C.sizeof(capi.DvrPeerSync),
           capi.DvrFrameParams.backgroundImage.offset, capi.DvrFrameParams.partialCullToBounds.offset,
           capi.DvrFrameBuffers.depthMirror.offset, C.sizeof(capi.DvrSlabExchange), capi.DvrSlabExchange.timing.offset,
           C.sizeof(capi.DvrSurfaceDesc), capi.DvrSurfaceDesc.objectToWorld.offset, C.sizeof(capi.DvrLight),
           C.sizeof(capi.DvrSceneParams), capi.DvrSceneParams.cullTriangleBackfaces.offset]
    assert got == want
    assert C.sizeof(capi.DvrCamera) == 4 + 16 + 12 * 6 + 8 and C.sizeof(capi.DvrFrameParams) == 96


def test_version():
    assert capi.version() == (0, 1)


def _cam_tuple(c):
    return (c.type, tuple(c.region), tuple(c.pos), tuple(c.dir), tuple(c.up), tuple(c.du), tuple(c.dv), tuple(c.p00),
            c.scaledAperture, c.aspect)


@pytest.mark.parametrize("args", [
    dict(pos=(1.0, 2.0, 3.0), direction=(-0.3, -0.5, -1.0), up=(0.0, 1.0, 0.0), fovy=1.0471976, aspect=16 / 9),
    dict(pos=(0.0, 0.0, 5.0), direction=(0.1, 0.0, -1.0), up=(0.2, 1.0, 0.1), fovy=0.6, aspect=0.75,
         focus_distance=3.5, aperture_radius=0.1, region=(0.1, 0.2, 0.9, 0.8)),
])
def test_camera_perspective_matches_oracle(args):
    assert _cam_tuple(capi.camera_perspective(**args)) == _cam_tuple(ob.camera_perspective(**args))


def test_camera_orthographic_matches_oracle():
    a = dict(pos=(4.0, -2.0, 1.0), direction=(-1.0, 0.4, 0.2), up=(0.0, 0.0, 1.0), height=2.5, aspect=1.3)
    assert _cam_tuple(capi.camera_orthographic(**a)) == _cam_tuple(ob.camera_orthographic(**a))


@pytest.mark.parametrize("case", range(len(TF_CASES)))
def test_tf_discretize_matches_oracle_and_reference(case):
    kw = TF_CASES[case]
    got = capi.tf_discretize(**kw)
    assert np.array_equal(got, ob.tf_discretize(which="cpu", **kw))
    if ob.have_ref_host():
        assert np.array_equal(got, ob.tf_discretize(which="ref", **kw))


def test_tf_discretize_rejects_bad_input():
    with pytest.raises(capi.DvrError):
        capi.tf_discretize(color=np.zeros((1, 4), np.float32))
    with pytest.raises(capi.DvrError):
        capi.tf_discretize(color=np.zeros((4, 2), np.float32))


def test_tsd_default_colormap_endpoints():
    cm = scenes.tsd_default_colormap(256)
    assert tuple(cm[0]) == (1.0, 0.0, 0.0, 0.0) and tuple(cm[-1]) == (0.0, 0.0, 1.0, 1.0)
    assert abs(cm[128, 1] - 1.0) < 0.01 and np.all(np.diff(cm[:, 3]) >= 0)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device error path")
def test_compute_entries_fail_loudly_without_a_gpu():
    assert capi.device_count() == 0
    vox = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(capi.DvrError) as e:
        capi.Field.create_structured(vox.ctypes.data, False, capi.DVR_FLOAT32, (4, 4, 4), (0, 0, 0), (1, 1, 1))
    assert e.value.code == capi.DVR_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    p = capi.frame_params(8, 8)
    cam = capi.camera_perspective((0, 0, 1), (0, 0, -1), (0, 1, 0), 1.0, 1.0)
    fb = capi.frame_buffers(1, 1)
    inst, n = capi.make_instances([])
    with pytest.raises(capi.DvrError) as e:
        capi.render(p, cam, inst, n, fb)
    assert e.value.code == capi.DVR_ERR_NO_DEVICE
    with pytest.raises(capi.DvrError) as e:
        capi.Surfaces.create([{"geometry": "sphere", "vertex.position": [(0.0, 0.0, 0.0)]}])
    assert e.value.code == capi.DVR_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_invalid_arguments_are_rejected():
    with pytest.raises(capi.DvrError) as e:
        capi.Field.create_structured(0, False, capi.DVR_FLOAT32, (4, 4, 4), (0, 0, 0), (1, 1, 1))
    assert e.value.code == capi.DVR_ERR_INVALID_ARGUMENT
    vox = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(capi.DvrError) as e:
        capi.Field.create_structured(vox.ctypes.data, False, 99, (4, 4, 4), (0, 0, 0), (1, 1, 1))
    assert e.value.code == capi.DVR_ERR_INVALID_ARGUMENT
    with pytest.raises(capi.DvrError):
        capi.Field.create_slab(vox.ctypes.data, False, capi.DVR_FLOAT32, (4, 4, 4), 3, 2, (0, 0, 0), (1, 1, 1))
    # the in-place refresh needs a field and data
    with pytest.raises(capi.DvrError) as e:
        capi.Field(None).update_structured(vox.ctypes.data, False, capi.DVR_FLOAT32, (0, 0, 0), (1, 1, 1))
    assert e.value.code == capi.DVR_ERR_INVALID_ARGUMENT and "dvr_field_update_structured" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "visrtx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                hit = re.search(r"liboracle|oracle_binding|oracle_[a-z]+\(|dvr_oracle|import oracle|oracle/|_ref/", text)
                assert hit is None, (os.path.join(dirpath, f), hit.group(0))
    import subprocess
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ref_" not in out


def test_post_pass_oracle_known_answers():
    """oracle/post_oracle.py (numpy restatement of TSD's render-pipeline passes) on hand-computed cases."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import post_oracle as PO
    # convertFloatColorBuffer: uint8(clamp(v) * 255) truncates
    assert PO.convert_float_color(np.array([[0.0, 1.0, 0.5, 2.0]], np.float32))[0] == (0 | 255 << 8 | 127 << 16 | 255 << 24)
    # shadePixel: 20 % of the pixel + 80 % of (1, .5, 0, 1), truncated to bytes
    assert PO.shade_pixel(np.array([0xFF000000], np.uint32))[0] == (204 | 102 << 8 | 0 << 16 | 255 << 24)
    assert PO.shade_pixel(np.array([0xFFFFFFFF], np.uint32))[0] == (255 | 153 << 8 | 50 << 16 | 255 << 24)  # 1 - .8f = .19999999
    # computeOutline: 3x3 counts 2..7 only; row 0 / column 0 never (unsigned wrap of `y - 1`)
    ids = np.zeros((5, 5), np.uint32)
    ids[1:4, 1:4] = 9
    col = np.zeros(25, np.uint32)
    out = PO.outline(col, ids.ravel(), 5, 5, 9).reshape(5, 5)
    assert (out[0] == 0).all() and (out[:, 0] == 0).all()
    assert out[2, 2] == 0  # centre: all 9 neighbours inside -> not an edge
    assert out[1, 1] != 0 and out[4, 4] == 0 and out[4, 3] != 0  # counts 4, 1 and 2
    # computeDepthImage: grey ramp, alpha 255; inf clamps to white
    v = PO.visualize_depth(np.array([0.0, 3.0, 6.0, np.inf], np.float32), 6.0)
    assert v.tolist() == [0xFF000000, 0xFF7F7F7F, 0xFFFFFFFF, 0xFFFFFFFF]
    # compositeFrame keeps the closer pixel
    c, d, i = PO.composite_depth(np.array([1, 2], np.uint32), np.array([0.5, 0.5], np.float32), np.array([8, 8], np.uint32),
                                 np.array([3, 4], np.uint32), np.array([0.25, 0.75], np.float32), np.array([9, 9], np.uint32), False)
    assert c.tolist() == [3, 2] and d.tolist() == [0.25, 0.5] and i.tolist() == [9, 8]


def test_bounds_screen_rect_is_conservative_for_random_cameras_and_boxes():
    """Every pixel with ANY jittered primary ray hitting the box must lie inside dvr_bounds_screen_rect (checked with
    numpy rays built from the same DvrCamera: 5 x 5 jitter positions per pixel incl. the pixel corners)."""
    rng = np.random.default_rng(42)
    W, Hh = 96, 64
    valid_seen = invalid_seen = shrunk = 0
    for trial in range(120):
        lo = rng.uniform(-2, 1, 3)
        hi = lo + rng.uniform(0.2, 2.5, 3)
        ctr = 0.5 * (lo + hi)
        pos = ctr + rng.normal(size=3) * rng.uniform(0.5, 8.0)
        look = ctr + rng.normal(size=3) * 0.8 - pos
        look /= np.linalg.norm(look)
        up = (0.0, 1.0, 0.0) if abs(look[1]) < 0.95 else (1.0, 0.0, 0.0)
        region = None if trial % 3 else tuple(sorted(rng.uniform(0, 1, 2))[i] for i in (0, 0, 1, 1))
        if region is not None:
            a, b = sorted(rng.uniform(0, 1, 2)); c, d = sorted(rng.uniform(0, 1, 2))
            region = (a, c, max(b, a + 0.05), max(d, c + 0.05))
        if trial % 4 == 3:
            cam = capi.camera_orthographic(tuple(pos), tuple(look), up, float(rng.uniform(1, 6)), W / Hh, region=region)
        else:
            cam = capi.camera_perspective(tuple(pos), tuple(look), up, float(rng.uniform(0.3, 1.6)), W / Hh, region=region)
        ok, (x0, y0, x1, y1) = capi.bounds_screen_rect(cam, lo, hi, W, Hh)
        if not ok:
            invalid_seen += 1
            assert (x0, y0, x1, y1) == (0, 0, W, Hh)
            continue
        valid_seen += 1
        shrunk += (x1 - x0) * (y1 - y0) < W * Hh
        hit = np.zeros((Hh, W), bool)
        reg = np.asarray(cam.region, np.float64)
        for jx in np.linspace(0.0, 0.999, 5):
            for jy in np.linspace(0.0, 0.999, 5):
                sx = (np.arange(W) + jx) / W
                sy = (np.arange(Hh) + jy) / Hh
                SX, SY = np.meshgrid(reg[0] + (reg[2] - reg[0]) * sx, reg[1] + (reg[3] - reg[1]) * sy)
                if cam.type == capi.DVR_CAMERA_PERSPECTIVE:
                    dirs = np.asarray(cam.p00)[None, None] + SX[..., None] * np.asarray(cam.du) + SY[..., None] * np.asarray(cam.dv)
                    org = np.broadcast_to(np.asarray(cam.pos, np.float64), dirs.shape)
                else:
                    org = np.asarray(cam.p00)[None, None] + SX[..., None] * np.asarray(cam.du) + SY[..., None] * np.asarray(cam.dv)
                    dirs = np.broadcast_to(np.asarray(cam.dir, np.float64), org.shape)
                with np.errstate(divide="ignore", invalid="ignore"):
                    t0 = (lo - org) / dirs
                    t1 = (hi - org) / dirs
                tn = np.minimum(t0, t1).max(-1)
                tf = np.maximum(t0, t1).min(-1)
                hit |= (tn < tf) & (tf >= 0)
        ys, xs = np.nonzero(hit)
        if len(xs):
            assert xs.min() >= x0 and xs.max() < x1 and ys.min() >= y0 and ys.max() < y1, (trial, (x0, y0, x1, y1))
    assert valid_seen > 40 and invalid_seen > 5 and shrunk > 20
    lens = capi.camera_perspective((0, 0, 5), (0, 0, -1), (0, 1, 0), 0.8, W / Hh, 4.0, 0.1)
    assert capi.bounds_screen_rect(lens, (-1, -1, -1), (1, 1, 1), W, Hh) == (False, (0, 0, W, Hh))
