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


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler gives the PODs (checked against ctypes mirrors; a mismatch would corrupt launches)
    assert C.sizeof(capi.DvrCamera) == 4 + 16 + 12 * 6 + 8
    assert C.sizeof(capi.DvrVolumeInstance) == 8 + 48 + 8
    assert C.sizeof(capi.DvrFrameBuffers) == 8 * 8
    assert C.sizeof(capi.DvrFrameParams) == 4 * 8 + 16 + 8 + 4 + 4 + 12 + 12
    assert C.sizeof(capi.DvrRenderStats) == 32


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
