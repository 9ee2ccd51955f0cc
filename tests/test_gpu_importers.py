"""GPU suite: the importers end to end (file -> tsd::import_volume equivalent -> ANARI field + volume -> frame) and
time-varying CUDA-pointer fields as the reference's animated_volume demo drives them
(tsd/apps/interactive/demos/animated_volume/SolverControls.cpp:159-214: in-place CUDA updates between map/unmap,
and swapping the field's `data` between two device arrays)."""
import ctypes as C
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
    # (3) a data array of another shape cannot refresh in place: field and volume are rebuilt
    m = 24
    small = scenes.marschner_lobb_np(m)
    buf_s = torch.from_numpy(small).cuda()
    other_s = d.new_array3d_device(buf_s.data_ptr(), A.FLOAT32, m, m, m)
    d.set(s.field, "data", A.ARRAY3D, other_s)
    d.commit(s.field)
    img3, n3 = shot(1)
    assert n3 == 0 and not np.array_equal(img3, img2)
    ref3 = AnariScene(m, 80, 60, "default", 0.5, vox=small)
    sp = 2.0 / (n - 1)
    ref3.d.set(ref3.field, "spacing", A.FLOAT32_VEC3, (sp, sp, sp))
    ref3.d.commit(ref3.field)
    ref3.d.set(ref3.volume, "unitDistance", A.FLOAT32, sp)
    ref3.d.commit(ref3.volume)
    for k in ("position", "direction", "up"):
        ref3.d.set(ref3.camera, k, A.FLOAT32_VEC3, getattr(s.pose, k))
    ref3.d.commit(ref3.camera)
    ref3.render()
    want3, _, _, _ = ref3.d.map_frame(ref3.frame, "channel.color")
    assert np.array_equal(img3, want3)
    # ... and back to the original shape
    d.set(s.field, "data", A.ARRAY3D, other)
    d.commit(s.field)
    img4, _ = shot(1)
    assert np.array_equal(img4, want0)
    assert not _errors(d), d.messages
    d.release(other)
    d.release(other_s)
    for x in (s, ref, first, ref3):
        x.close()


@pytest.mark.parametrize("elem", ["f32_device", "f32_host", "ufixed8_host", "f16_device"])
def test_field_update_in_place_equals_fresh_field(elem):
    """dvr_field_update_structured + dvr_volume_update == destroying and re-creating field and volume: same frame,
    same macrocell ranges, for device and host data, including a change of origin / spacing."""
    import torch
    n = 40
    a0 = scenes.marschner_lobb_np(n)
    a1 = np.ascontiguousarray(a0[::-1, ::-1, :])
    dt = capi.DVR_FLOAT32
    if elem == "ufixed8_host":
        a0, a1 = ((a * 255).astype(np.uint8) for a in (a0, a1))
        dt = capi.DVR_UFIXED8
    elif elem == "f16_device":
        a0, a1 = (a.astype(np.float16) for a in (a0, a1))
        dt = capi.DVR_FLOAT16
    on_dev = elem.endswith("device")
    keep = []

    def ptr(a):
        if on_dev:
            t = torch.from_numpy(a).cuda()
            keep.append(t)
            return t.data_ptr()
        return a.ctypes.data

    sp0, sp1 = 2.0 / (n - 1), 1.5 / (n - 1)
    tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256))
    f = capi.Field.create_structured(ptr(a0), on_dev, dt, (n, n, n), (-1, -1, -1), (sp0,) * 3)
    v = capi.Volume.create(f, tf, (0.0, 1.0), sp0, 3)
    f.update_structured(ptr(a1), on_dev, dt, (-0.75, -0.75, -0.75), (sp1,) * 3)
    v.update(tf, (0.0, 1.0), sp0, 3)
    g = capi.Field.create_structured(ptr(a1), on_dev, dt, (n, n, n), (-0.75, -0.75, -0.75), (sp1,) * 3)
    w = capi.Volume.create(g, tf, (0.0, 1.0), sp0, 3)
    assert f.bounds() == g.bounds() and f.step_size() == g.step_size()

    def ranges(fld):
        (gx, gy, gz), p = fld.macrocells()
        r = torch.empty((gz, gy, gx, 2), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(r.data_ptr()), C.c_void_p(p), C.c_size_t(r.numel() * 4), C.c_int(3))
        return r.cpu().numpy()

    assert np.array_equal(ranges(f), ranges(g))
    pose = scenes.orbit_camera((-1, -1, -1), (1, 1, 1), 96, 64)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    npx = 96 * 64
    imgs = []
    for vol in (v, w):
        inst, ni = capi.make_instances([vol], None, [0])
        accum = torch.zeros((npx, 4), dtype=torch.float32, device="cuda")
        color = torch.zeros(npx, dtype=torch.int32, device="cuda")
        fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr())
        params = capi.frame_params(96, 64, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_RAYCAST, background=(0.1, 0.1, 0.1, 1.0),
                                   volume_sampling_rate=0.5)
        capi.render(params, cam, inst, ni, fb)
        torch.cuda.synchronize()
        imgs.append(color.cpu().numpy())
    assert np.array_equal(imgs[0], imgs[1])
    assert len(np.unique(imgs[0])) > 50
    # a slab, a NanoVDB grid or another element type do not update in place
    with pytest.raises(capi.DvrError):
        f.update_structured(ptr(a1), on_dev, capi.DVR_FIXED16, (0, 0, 0), (1, 1, 1))
    for o in (v, w, f, g):
        o.destroy()
