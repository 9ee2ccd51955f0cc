"""GPU suite: the importers end to end (file -> tsd::import_volume equivalent -> ANARI field + volume -> frame) and
time-varying CUDA-pointer fields as the reference's animated_volume demo drives them
(tsd/apps/interactive/demos/animated_volume/SolverControls.cpp:159-214: in-place CUDA updates between map/unmap,
and swapping the field's `data` between two device arrays)."""
import os

import numpy as np
import pytest

import dvr_harness as H
from test_gpu_anari import AnariScene, _errors
from visrtx_b200 import anari as A
from visrtx_b200 import capi, importers as I, nvdb_writer, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _frame_for(d, volume, bounds, w=96, h=96, renderer="raycast"):
    world = d.new("World")
    vols = d.new_object_array([volume], A.VOLUME)
    d.set(world, "volume", A.ARRAY1D, vols)
    d.commit(world)
    lo, hi = bounds
    pose = scenes.orbit_camera(lo, hi, w, h)
    cam = d.new("Camera", "perspective")
    for k, t, v in (("position", A.FLOAT32_VEC3, pose.position), ("direction", A.FLOAT32_VEC3, pose.direction),
                    ("up", A.FLOAT32_VEC3, pose.up), ("fovy", A.FLOAT32, pose.fovy), ("aspect", A.FLOAT32, pose.aspect)):
        d.set(cam, k, t, v)
    d.commit(cam)
    ren = d.new("Renderer", renderer)
    d.set(ren, "background", A.FLOAT32_VEC4, (0.1, 0.1, 0.1, 1.0))
    d.set(ren, "volumeSamplingRate", A.FLOAT32, 0.5)
    d.commit(ren)
    frame = d.new("Frame")
    d.set(frame, "size", A.UINT32_VEC2, (w, h))
    d.set(frame, "channel.color", A.DATA_TYPE, A.UFIXED8_RGBA_SRGB)
    d.set(frame, "renderer", A.RENDERER, ren)
    d.set(frame, "camera", A.CAMERA, cam)
    d.set(frame, "world", A.WORLD, world)
    d.commit(frame)
    return frame, pose


def test_imported_raw_volume_renders_like_the_same_data_through_the_cabi(tmp_path):
    rng = np.random.default_rng(8)
    vox = (scenes.blobs_np(32) * 60000).astype(np.uint16)
    p = tmp_path / "blobs_32x32x32_uint16.raw"
    vox.tofile(p)
    d = A.Device()
    volume, field, vf = I.import_volume(d, str(p))
    assert vf.data_type == capi.DVR_UFIXED16 and vf.dims == (32, 32, 32)
    lo, hi = (0.0,) * 3, (31.0,) * 3
    frame, pose = _frame_for(d, volume, (lo, hi))
    d.render(frame)
    d.wait(frame)
    color, w, h, _ = d.map_frame(frame, "channel.color")
    assert not _errors(d), d.messages
    # the same scene assembled by hand: import_volume uses the default colour map and valueRange = data range
    v = H.VolumeDesc(vox, data_type=capi.DVR_UFIXED16, origin=lo, spacing=(1.0,) * 3, value_range=vf.value_range,
                     tf=capi.tf_discretize(color=scenes.tsd_default_colormap(256), value_range=vf.value_range))
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    scene = H.SceneDesc([v], 96, 96, cam, volume_sampling_rate=0.5)
    ref = H.render_cuda(scene)
    assert np.array_equal(color, ref["color"])
    d.close()


@pytest.mark.parametrize("name", ["fog_float_zip.nvdb", "fog_fp8_zip.nvdb", "fog_fpn_zip.nvdb"])
def test_imported_nvdb_file_renders_like_the_grid_it_holds(name, tmp_path):
    fix = np.load(os.path.join(GOLD, "import_fixtures.npz"))
    p = tmp_path / name
    fix["file/" + name].tofile(p)
    d = A.Device()
    volume, field, vf = I.import_volume(d, str(p))
    assert vf.kind == I.NANOVDB
    wb = vf.data[560:608].view(np.float64)
    frame, pose = _frame_for(d, volume, (wb[:3].astype(np.float32), wb[3:].astype(np.float32)), renderer="default")
    for _ in range(2):
        d.render(frame)
        d.wait(frame)
    color, _, _, _ = d.map_frame(frame, "channel.color")
    assert not _errors(d), d.messages
    v = H.VolumeDesc(np.zeros((1, 1, 1), np.float32), nvdb=vf.data, value_range=vf.value_range,
                     tf=capi.tf_discretize(color=scenes.tsd_default_colormap(256), value_range=vf.value_range))
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    scene = H.SceneDesc([v], 96, 96, cam, volume_sampling_rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT)
    ref = H.render_cuda(scene, frames=2)
    assert np.array_equal(color, ref["color"])
    want = H.render_oracle(scene, frames=2)
    dd = np.abs(H.unpack_rgba8(color) - H.unpack_rgba8(want["color"])).max(axis=-1)
    assert (dd <= 2).mean() >= 0.999
    d.close()


def test_time_varying_cuda_pointer_field():
    """In-situ style updates: the simulation writes into the CUDA array the field shares (map/unmap brackets the
    write), or flips the field between two device arrays; each update restarts accumulation and shows the new data."""
    import torch
    n = 40
    a0 = scenes.marschner_lobb_np(n)
    a1 = np.ascontiguousarray(a0[::-1, :, ::-1])  # a different field of the same size
    buf_a = torch.from_numpy(a0).cuda()
    buf_b = torch.from_numpy(a1).cuda()
    s = AnariScene(n, 80, 60, "default", 0.5, vox=a0, device_ptr=buf_a.data_ptr())
    d = s.d

    def shot(frames=1):
        for _ in range(frames):
            s.render()
        c, _, _, _ = d.map_frame(s.frame, "channel.color")
        return np.array(c, copy=True), d.get_property(s.frame, "numSamples", A.INT32)

    img0, n0 = shot(3)
    assert n0 == 2  # numSamples = samples accumulated before the last frame (Frame.cu frameID)
    # (1) in-place device write between map/unmap (SolverControls.cpp:196-203)
    assert d.map_array(s.data) == buf_a.data_ptr()  # a shared CUDA array maps to its own device pointer
    buf_a.copy_(buf_b)
    torch.cuda.synchronize()
    d.unmap_array(s.data)
    img1, n1 = shot(1)
    assert n1 == 0 and not np.array_equal(img1, img0)  # accumulation restarted on the new data
    ref = AnariScene(n, 80, 60, "default", 0.5, vox=a1)
    ref.render()
    want, _, _, _ = ref.d.map_frame(ref.frame, "channel.color")
    assert np.array_equal(img1, want)
    # (2) ping-pong between two device arrays (SolverControls.cpp:204-207)
    buf_c = torch.from_numpy(a0).cuda()
    other = d.new_array3d_device(buf_c.data_ptr(), A.FLOAT32, n, n, n)
    d.set(s.field, "data", A.ARRAY3D, other)
    d.commit(s.field)
    img2, n2 = shot(1)
    assert n2 == 0
    first = AnariScene(n, 80, 60, "default", 0.5, vox=a0)
    first.render()
    want0, _, _, _ = first.d.map_frame(first.frame, "channel.color")
    assert np.array_equal(img2, want0)
    assert not _errors(d), d.messages
    d.release(other)
    for x in (s, ref, first):
        x.close()
