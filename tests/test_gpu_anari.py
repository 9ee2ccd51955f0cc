"""GPU suite, ANARI boundary: the same scenes driven through the ANARI C API (as an application would)
must give exactly the frames of the C-ABI path (which the parity tests pin to the oracle), and the
frame/renderer/array semantics of the reference must hold: deferred commits, accumulation reset on any
finalisation, sampleLimit, duration/numSamples/nextFrameReset, *CUDA channel maps, CUDA-pointer arrays,
completion callback, world bounds, instance transforms."""
import ctypes as C
import math

import numpy as np
import pytest

import dvr_harness as H
from visrtx_b200 import anari as A
from visrtx_b200 import capi, scenes

pytestmark = pytest.mark.gpu


class AnariScene:
    def __init__(self, n=48, w=96, h=96, renderer="raycast", rate=0.5, color_type=A.UFIXED8_RGBA_SRGB,
                 channels=("depth", "objectId", "instanceId", "primitiveId"), vox=None, elem=A.FLOAT32, device_ptr=None,
                 gpus=None, multi_gpu_mode=None):
        self.d = d = A.Device()
        if gpus is not None:  # multi-GPU device: "cudaDevices" (display GPU first) + "multiGpuMode"
            d.set(d.handle, "cudaDevices", A.STRING, ",".join(str(g) for g in gpus))
            if multi_gpu_mode:
                d.set(d.handle, "multiGpuMode", A.STRING, multi_gpu_mode)
            d.commit(d.handle)
        self.n, self.w, self.h = n, w, h
        sp = 2.0 / (n - 1)
        self.vox = scenes.marschner_lobb_np(n) if vox is None else vox
        if device_ptr is not None:
            self.data = d.new_array3d_device(device_ptr, elem, n, n, n)
        else:
            self.data = d.new_array3d(self.vox, elem)
        self.field = d.new("SpatialField", "structuredRegular")
        d.set(self.field, "data", A.ARRAY3D, self.data)
        d.set(self.field, "origin", A.FLOAT32_VEC3, (-1, -1, -1))
        d.set(self.field, "spacing", A.FLOAT32_VEC3, (sp, sp, sp))
        d.commit(self.field)
        self.volume = d.new("Volume", "transferFunction1D")
        self.cmap = scenes.tsd_default_colormap(256)
        self.color = d.new_array1d(self.cmap, A.FLOAT32_VEC4)
        d.set(self.volume, "color", A.ARRAY1D, self.color)
        d.set(self.volume, "value", A.SPATIAL_FIELD, self.field)
        d.set(self.volume, "valueRange", A.FLOAT32_BOX1, (0.0, 1.0))
        d.set(self.volume, "unitDistance", A.FLOAT32, sp)
        d.set(self.volume, "id", A.UINT32, 7)
        d.commit(self.volume)
        self.world = d.new("World")
        self.vols = d.new_object_array([self.volume], A.VOLUME)
        d.set(self.world, "volume", A.ARRAY1D, self.vols)
        d.commit(self.world)
        lo = np.float32(-1.0)
        hi = lo + (np.float32(n) - np.float32(1.0)) * np.float32(sp)  # field bounds as the device computes them
        self.pose = scenes.orbit_camera((lo,) * 3, (hi,) * 3, w, h)
        self.camera = d.new("Camera", "perspective")
        d.set(self.camera, "position", A.FLOAT32_VEC3, self.pose.position)
        d.set(self.camera, "direction", A.FLOAT32_VEC3, self.pose.direction)
        d.set(self.camera, "up", A.FLOAT32_VEC3, self.pose.up)
        d.set(self.camera, "fovy", A.FLOAT32, self.pose.fovy)
        d.set(self.camera, "aspect", A.FLOAT32, self.pose.aspect)
        d.commit(self.camera)
        self.renderer = d.new("Renderer", renderer)
        d.set(self.renderer, "background", A.FLOAT32_VEC4, (0.1, 0.1, 0.1, 1.0))
        d.set(self.renderer, "volumeSamplingRate", A.FLOAT32, rate)
        d.commit(self.renderer)
        self.frame = d.new("Frame")
        d.set(self.frame, "size", A.UINT32_VEC2, (w, h))
        d.set(self.frame, "channel.color", A.DATA_TYPE, color_type)
        for ch in channels:
            d.set(self.frame, "channel." + ch, A.DATA_TYPE, A.FLOAT32 if ch == "depth" else A.UINT32)
        d.set(self.frame, "renderer", A.RENDERER, self.renderer)
        d.set(self.frame, "camera", A.CAMERA, self.camera)
        d.set(self.frame, "world", A.WORLD, self.world)
        d.commit(self.frame)

    def render(self):
        self.d.render(self.frame)
        self.d.wait(self.frame)

    def close(self):
        d = self.d
        for o in (self.frame, self.renderer, self.camera, self.world, self.vols, self.volume, self.color, self.field,
                  self.data):
            d.release(o)
        d.close()


def _errors(dev):
    return [m for m in dev.messages if m[0] <= A.SEVERITY_ERROR]


def test_anari_frame_equals_cabi_frame_and_oracle():
    s = AnariScene(48, 96, 96, "raycast", 0.5)
    s.render()
    color, w, h, t = s.d.map_frame(s.frame, "channel.color")
    depth, _, _, td = s.d.map_frame(s.frame, "channel.depth")
    obj_id, _, _, _ = s.d.map_frame(s.frame, "channel.objectId")
    inst_id, _, _, _ = s.d.map_frame(s.frame, "channel.instanceId")
    assert (w, h, t, td) == (96, 96, A.UFIXED8_RGBA_SRGB, A.FLOAT32)
    assert not _errors(s.d), s.d.messages
    scene = H.default_scene(48, 96, 96, rate=0.5)
    scene.volumes[0].inst_id = 0xFFFFFFFF  # world-level volumes live in the zero instance (id ~0u)
    ref = H.render_cuda(scene)
    assert np.array_equal(color, ref["color"])
    assert np.array_equal(depth, ref["depth"])
    assert np.array_equal(obj_id, ref["objId"]) and np.array_equal(inst_id, ref["instId"])
    want = H.render_oracle(scene)
    d = np.abs(H.unpack_rgba8(color) - H.unpack_rgba8(want["color"])).max(axis=-1)
    assert (d <= 2).mean() >= 0.999
    s.close()


def test_background_image_through_anari_equals_cabi_and_reference():
    """Renderer parameter "background" as an Array2D (Renderer.cpp:154,175-198): sampled per pixel-sample at its screen
    coordinate.  ANARI frame == C-ABI frame (same texture path) and within north_star's tolerance of O-gpu, which runs
    the reference's own getBackground; replacing the array's contents (map/unmap) re-finalises the renderer."""
    import oracle_binding as ob
    zoo_scene, frames, _ = H.scene_zoo()["bgimage_float3_default_spp2_f2"]
    img = zoo_scene.background_image[0]
    s = AnariScene(32, 96, 96, "default", 0.5, color_type=A.FLOAT32_VEC4, channels=("depth", "objectId"),
                   vox=scenes.blobs_np(32))
    d = s.d
    d.set(s.volume, "unitDistance", A.FLOAT32, 0.6)
    d.commit(s.volume)
    bg = d.new_array2d(img, A.FLOAT32_VEC3)
    d.set(s.renderer, "background", A.ARRAY2D, bg)
    d.set(s.renderer, "pixelSamples", A.INT32, 2)
    d.commit(s.renderer)
    for _ in range(frames):
        s.render()
    color, w, h, t = d.map_frame(s.frame, "channel.color")
    assert not _errors(d), d.messages
    assert not [m for m in d.messages if "background" in m[2]], d.messages  # no "ignored" warning any more
    zoo_scene.volumes[0].inst_id = 0xFFFFFFFF
    ref = H.render_cuda(zoo_scene, frames=frames)
    assert np.array_equal(color.reshape(-1, 4), ref["color"])
    if ob.have_ref_gpu():
        want = H.render_refgpu(zoo_scene, frames=frames)
        dd = np.abs(color.reshape(-1, 4) - want["color"]).max(axis=-1) * 255.0
        assert dd.max() <= 2.0
    # the image is really used: a constant-colour render differs in the missed pixels
    d.unset(s.renderer, "background")
    d.set(s.renderer, "background", A.FLOAT32_VEC4, (0.1, 0.1, 0.1, 1.0))
    d.commit(s.renderer)
    s.render()
    flat, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert np.abs(flat.reshape(-1, 4) - color.reshape(-1, 4)).max() > 0.2
    d.release(bg)
    s.close()


def test_accumulation_reset_semantics_and_frame_properties():
    s = AnariScene(32, 64, 64, "default", 0.5, color_type=A.FLOAT32_VEC4, channels=())
    d = s.d
    assert d.get_property(s.frame, "nextFrameReset", A.BOOL) == 1
    s.render()
    assert d.get_property(s.frame, "numSamples", A.INT32) == 0
    first, _, _, _ = d.map_frame(s.frame, "channel.color")
    s.render()
    s.render()
    assert d.get_property(s.frame, "numSamples", A.INT32) == 2  # frameID counts accumulated samples
    assert d.get_property(s.frame, "nextFrameReset", A.BOOL) == 0
    acc, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert not np.array_equal(first, acc)
    dur = d.get_property(s.frame, "duration", A.FLOAT32)
    assert 0.0 < dur < 1.0
    # any commit restarts accumulation (Frame.cu:574-588) ...
    d.set(s.camera, "fovy", A.FLOAT32, s.pose.fovy)
    d.commit(s.camera)
    assert d.get_property(s.frame, "nextFrameReset", A.BOOL) == 1
    s.render()
    assert d.get_property(s.frame, "numSamples", A.INT32) == 0
    again, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert np.array_equal(first, again)  # same frameID 0 => same jitter => identical frame
    # ... and so does an array edit through map/unmap (change observers, Array.cpp:152-162)
    s.render()
    p = d.map_array(s.color)
    cm = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(256, 4))
    cm[:, 0] = 0.0
    d.unmap_array(s.color)
    assert d.get_property(s.frame, "nextFrameReset", A.BOOL) == 1
    s.render()
    edited, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert d.get_property(s.frame, "numSamples", A.INT32) == 0
    assert not np.array_equal(first, edited)
    assert not _errors(d)
    s.close()


def test_raycast_is_single_shot_and_default_honours_sample_limit():
    s = AnariScene(24, 48, 48, "raycast", 0.5, channels=())
    for _ in range(4):
        s.render()
    assert s.d.get_property(s.frame, "numSamples", A.INT32) == 1  # Raycast.cpp:48 sampleLimit = 1 (quirk Q9)
    s.close()
    s = AnariScene(24, 48, 48, "default", 0.5, channels=())
    s.d.set(s.renderer, "sampleLimit", A.INT32, 3)
    s.d.set(s.renderer, "pixelSamples", A.INT32, 2)
    s.d.commit(s.renderer)
    for _ in range(6):
        s.render()
    assert s.d.get_property(s.frame, "numSamples", A.INT32) == 4  # 0,2,4 then >= 3 stops
    s.close()


def test_cuda_channel_maps_and_cuda_array_input():
    import torch
    vox = scenes.marschner_lobb_np(40)
    dev_vox = torch.from_numpy(vox).cuda()
    s = AnariScene(40, 80, 60, "raycast", 0.5, vox=vox, device_ptr=dev_vox.data_ptr())
    s.render()
    host, w, h, _ = s.d.map_frame(s.frame, "channel.color")
    ptr, w2, h2, t = s.d.map_frame(s.frame, "channel.colorCUDA")
    assert (w2, h2, t) == (80, 60, A.UFIXED8_RGBA_SRGB) and isinstance(ptr, int) and ptr != 0
    out = torch.empty(w * h, dtype=torch.int32, device="cuda")
    C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), C.c_size_t(w * h * 4), C.c_int(3))
    assert np.array_equal(out.cpu().numpy().view(np.uint32), host)
    dptr, _, _, td = s.d.map_frame(s.frame, "channel.depthCUDA")
    assert td == A.FLOAT32 and dptr
    # deprecated alias still works, with a warning (Frame.cu:371-411)
    gptr, _, _, _ = s.d.map_frame(s.frame, "channel.colorGPU")
    assert gptr == ptr and any("deprecated" in m[2] for m in s.d.messages)
    # unknown / disabled channels map to NULL with type UNKNOWN
    none, _, _, tn = s.d.map_frame(s.frame, "channel.albedo")
    assert none is None and tn == A.UNKNOWN
    # same image as the host-array upload
    s2 = AnariScene(40, 80, 60, "raycast", 0.5, vox=vox)
    s2.render()
    host2, _, _, _ = s2.d.map_frame(s2.frame, "channel.color")
    assert np.array_equal(host, host2)
    s.close()
    s2.close()


def test_frame_completion_callback_and_async_render():
    s = AnariScene(32, 64, 64, "default", 0.5, channels=())
    hits = []
    cb = A.FrameCompletionCallback(lambda user, dev, frame: hits.append(frame))
    s.d.set(s.frame, "frameCompletionCallback", A.FRAME_COMPLETION_CALLBACK, C.cast(cb, C.c_void_p).value)
    s.d.commit(s.frame)
    s.d.render(s.frame)  # returns after enqueue
    assert s.d.wait(s.frame) == 1 and s.d.is_ready(s.frame)
    assert hits == [s.frame]
    s.close()


def test_world_bounds_and_instance_transform():
    s = AnariScene(32, 96, 96, "raycast", 0.5)
    d = s.d
    b = d.get_property(s.world, "bounds", A.FLOAT32_BOX3)
    np.testing.assert_allclose(b, (-1, -1, -1, 1, 1, 1), atol=1e-6)
    s.render()
    base, _, _, _ = d.map_frame(s.frame, "channel.color")
    # move the same volume into a translated instance: rendering it from a camera translated by the same
    # offset reproduces the image; instance id lands in the instanceId channel
    group = d.new("Group")
    d.set(group, "volume", A.ARRAY1D, s.vols)
    d.commit(group)
    inst = d.new("Instance", "transform")
    d.set(inst, "group", A.GROUP, group)
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = (0.5, -0.25, 2.0)  # column-major: translation in the last column
    d.set(inst, "transform", A.FLOAT32_MAT4, m)
    d.set(inst, "id", A.UINT32, 42)
    d.commit(inst)
    insts = d.new_object_array([inst], A.INSTANCE)
    d.unset(s.world, "volume")
    d.set(s.world, "instance", A.ARRAY1D, insts)
    d.commit(s.world)
    pos = tuple(np.float32(p) + np.float32(o) for p, o in zip(s.pose.position, (0.5, -0.25, 2.0)))
    d.set(s.camera, "position", A.FLOAT32_VEC3, pos)
    d.commit(s.camera)
    b = d.get_property(s.world, "bounds", A.FLOAT32_BOX3)
    np.testing.assert_allclose(b, (-0.5, -1.25, 1, 1.5, 0.75, 3), atol=1e-5)
    s.render()
    moved, _, _, _ = d.map_frame(s.frame, "channel.color")
    inst_id, _, _, _ = d.map_frame(s.frame, "channel.instanceId")
    dd = np.abs(H.unpack_rgba8(moved) - H.unpack_rgba8(base)).max(axis=-1)
    assert (dd <= 1).mean() > 0.999
    depth, _, _, _ = d.map_frame(s.frame, "channel.depth")
    assert set(np.unique(inst_id[depth < 1e29])) == {42}
    assert not _errors(d)
    for o in (insts, inst, group):
        d.release(o)
    s.close()


def test_checkerboarding_and_invalid_objects_are_skipped_with_warnings():
    s = AnariScene(24, 50, 38, "default", 0.5, channels=())
    d = s.d
    d.set(s.renderer, "checkerboarding", A.BOOL, 1)
    d.commit(s.renderer)
    for _ in range(5):
        s.render()
    assert d.get_property(s.frame, "numSamples", A.INT32) == 1  # frameID advances after 4 checkerboard passes
    img, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert len(np.unique(img)) > 10
    # a volume without a field is skipped, the frame still renders
    bad = d.new("Volume", "transferFunction1D")
    d.commit(bad)
    both = d.new_object_array([s.volume, bad], A.VOLUME)
    d.set(s.world, "volume", A.ARRAY1D, both)
    d.commit(s.world)
    s.render()
    assert any("missing parameter 'value'" in m[2] for m in d.messages)
    unknown = d.new("SpatialField", "amr")
    d.commit(unknown)
    s.render()
    assert any("unknown spatial field subtype" in m[2] for m in d.messages)
    for o in (both, bad, unknown):
        d.release(o)
    s.close()


def test_field_value_range_property_and_float16_extension():
    vox = (scenes.blobs_np(32) * 3.0 - 1.0).astype(np.float16)
    s = AnariScene(32, 32, 32, "raycast", 0.5, vox=vox, elem=A.FLOAT16, channels=())
    r = s.d.get_property(s.field, "valueRange", A.FLOAT32_BOX1)
    assert r == (float(vox.min()), float(vox.max()))
    s.render()
    assert not _errors(s.d)
    s.close()


def test_nanovdb_field_through_anari_matches_cabi():
    """anariNewSpatialField("nanovdb") with a UINT8 Array1D holding one serialized grid (NvdbRegularField)."""
    import os
    fog = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvdb_fog_spheres.npz"))
    blob = np.ascontiguousarray(fog["r20"])
    scene, frames, _ = H.scene_zoo()["nvdb_fog_r20"]
    scene.volumes[0].inst_id = 0xFFFFFFFF
    want = H.render_cuda(scene, frames=1)
    d = A.Device()
    data = d.new_array1d(blob, A.UINT8)
    field = d.new("SpatialField", "nanovdb")
    d.set(field, "data", A.ARRAY1D, data)
    d.commit(field)
    vol = d.new("Volume", "transferFunction1D")
    col = d.new_array1d(scenes.tsd_default_colormap(256), A.FLOAT32_VEC4)
    d.set(vol, "color", A.ARRAY1D, col)
    d.set(vol, "value", A.SPATIAL_FIELD, field)
    d.set(vol, "unitDistance", A.FLOAT32, 4.0)
    d.set(vol, "id", A.UINT32, 31)
    d.commit(vol)
    world = d.new("World")
    vols = d.new_object_array([vol], A.VOLUME)
    d.set(world, "volume", A.ARRAY1D, vols)
    d.commit(world)
    b = d.get_property(world, "bounds", A.FLOAT32_BOX3)
    assert b == (-20.0, -20.0, -20.0, 21.0, 21.0, 21.0)
    pose = scenes.orbit_camera(b[:3], b[3:], scene.width, scene.height, dist_scale=1.0)
    cam = d.new("Camera", "perspective")
    for k, t, v in (("position", A.FLOAT32_VEC3, pose.position), ("direction", A.FLOAT32_VEC3, pose.direction),
                    ("up", A.FLOAT32_VEC3, pose.up), ("fovy", A.FLOAT32, pose.fovy), ("aspect", A.FLOAT32, pose.aspect)):
        d.set(cam, k, t, v)
    d.commit(cam)
    ren = d.new("Renderer", "default")
    d.set(ren, "background", A.FLOAT32_VEC4, (0.1, 0.1, 0.1, 1.0))
    d.set(ren, "volumeSamplingRate", A.FLOAT32, 0.5)
    d.commit(ren)
    frame = d.new("Frame")
    d.set(frame, "size", A.UINT32_VEC2, (scene.width, scene.height))
    d.set(frame, "channel.color", A.DATA_TYPE, A.UFIXED8_RGBA_SRGB)
    d.set(frame, "channel.depth", A.DATA_TYPE, A.FLOAT32)
    for k, t, o in (("renderer", A.RENDERER, ren), ("camera", A.CAMERA, cam), ("world", A.WORLD, world)):
        d.set(frame, k, t, o)
    d.commit(frame)
    d.render(frame)
    d.wait(frame)
    color, _, _, _ = d.map_frame(frame, "channel.color")
    assert not _errors(d), d.messages
    assert np.array_equal(color, want["color"])
    # a grid type outside Float / Fp4 / Fp8 / Fp16 / FpN is refused with a warning, not a crash
    bad = blob.copy()
    bad[636:640] = np.frombuffer(np.uint32(2).tobytes(), np.uint8)  # GridType::Double
    f2 = d.new("SpatialField", "nanovdb")
    d.set(f2, "data", A.ARRAY1D, d.new_array1d(bad, A.UINT8))
    d.commit(f2)
    d.render(frame)
    assert any("unsupported GridType" in m[2] for m in d.messages)
    d.close()


def test_host_colour_streaming_equals_copy_path():
    """After the first host map of channel.color the device streams the encoded colour into pinned host memory
    during the launch (DvrFrameBuffers::outColorMirror).  Streamed frames must equal the device buffer and the
    C-ABI renders bit for bit, for every colour format, through resets and sampleLimit stops."""
    import torch
    for ctype, fmt in ((A.UFIXED8_RGBA_SRGB, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB), (A.FLOAT32_VEC4, capi.DVR_FORMAT_FLOAT32_VEC4),
                       (A.UFIXED8_VEC4, capi.DVR_FORMAT_UFIXED8_VEC4)):
        s = AnariScene(40, 72, 56, "default", 0.5, color_type=ctype, channels=("depth",))
        d = s.d
        scene = H.default_scene(40, 72, 56, rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT, fmt=fmt)
        scene.volumes[0].inst_id = 0xFFFFFFFF
        for frames in (1, 2, 3):  # frame 1: copy path; frames 2, 3: streamed
            s.render()
            host, w, h, _ = d.map_frame(s.frame, "channel.color")
            host = np.array(host, copy=True)
            ptr, _, _, _ = d.map_frame(s.frame, "channel.colorCUDA")
            nbytes = w * h * (16 if fmt == capi.DVR_FORMAT_FLOAT32_VEC4 else 4)
            dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(dev.data_ptr()), C.c_void_p(ptr), C.c_size_t(nbytes), C.c_int(3))
            assert np.array_equal(np.ascontiguousarray(host).view(np.uint8).ravel(), dev.cpu().numpy()), (fmt, frames)
            ref = H.render_cuda(scene, frames=frames)
            assert np.array_equal(host.reshape(ref["color"].shape), ref["color"]), (fmt, frames)
        # a parameter change resets accumulation: the streamed frame is frame 0 again
        d.set(s.renderer, "background", A.FLOAT32_VEC4, (0.3, 0.2, 0.1, 1.0))
        d.commit(s.renderer)
        s.render()
        host, _, _, _ = d.map_frame(s.frame, "channel.color")
        scene.background = (0.3, 0.2, 0.1, 1.0)
        ref = H.render_cuda(scene, frames=1)
        assert np.array_equal(np.asarray(host).reshape(ref["color"].shape), ref["color"])
        # sampleLimit reached: no launch, the mapped image stays the last one
        d.set(s.renderer, "sampleLimit", A.INT32, 1)
        d.commit(s.renderer)
        s.render()
        s.render()  # frameID 0 and 1 are rendered; from then on frameID >= sampleLimit stops (Frame.cu:251-253)
        last = np.array(d.map_frame(s.frame, "channel.color")[0], copy=True)
        s.render()
        again = np.array(d.map_frame(s.frame, "channel.color")[0], copy=True)
        assert np.array_equal(last, again)
        assert not _errors(d), d.messages
        s.close()


def test_array1d_region_parameters_begin_end():
    """Array1D honours `begin` / `end` (array/Array1D.cpp:43-66): consumers see elements [begin, end) only — the
    transfer function's colour control points and the world's volume list."""
    n = 32
    cmap = np.array([[9, 9, 9, 9], [1, 0, 0, 0.0], [0, 1, 0, 0.5], [0, 0, 1, 1.0], [7, 7, 7, 7], [5, 5, 5, 5]], np.float32)
    s = AnariScene(n, 64, 48, "raycast", 0.5)
    d = s.d
    arr = d.new_array1d(cmap, A.FLOAT32_VEC4)
    d.set(arr, "begin", A.UINT64, 1)
    d.set(arr, "end", A.UINT64, 4)
    d.commit(arr)
    d.set(s.volume, "color", A.ARRAY1D, arr)
    d.commit(s.volume)
    s.render()
    got, _, _, _ = d.map_frame(s.frame, "channel.color")
    ref = AnariScene(n, 64, 48, "raycast", 0.5)
    mid = ref.d.new_array1d(np.ascontiguousarray(cmap[1:4]), A.FLOAT32_VEC4)
    ref.d.set(ref.volume, "color", A.ARRAY1D, mid)
    ref.d.commit(ref.volume)
    ref.render()
    want, _, _, _ = ref.d.map_frame(ref.frame, "channel.color")
    assert np.array_equal(got, want)
    # swapped bounds warn and are swapped; a later change of the window re-finalises the volume (accumulation resets)
    d.set(arr, "begin", A.UINT64, 4)
    d.set(arr, "end", A.UINT64, 1)
    d.commit(arr)
    s.render()
    again, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert np.array_equal(again, want) and any("swapping" in m[2] for m in d.messages)
    d.set(arr, "begin", A.UINT64, 2)
    d.set(arr, "end", A.UINT64, 5)
    d.commit(arr)
    s.render()
    other, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert not np.array_equal(other, want) and d.get_property(s.frame, "numSamples", A.INT32) == 0
    # object arrays: the world renders only the volumes inside the window
    two = d.new_object_array([s.volume, s.volume], A.VOLUME)
    d.set(two, "begin", A.UINT64, 1)
    d.commit(two)
    d.set(s.world, "volume", A.ARRAY1D, two)
    d.commit(s.world)
    s.render()
    one_vol, _, _, _ = d.map_frame(s.frame, "channel.color")
    assert np.array_equal(one_vol, other)  # one instance of the volume, not two overlapping ones
    assert not _errors(d), d.messages
    s.close()
    ref.close()
