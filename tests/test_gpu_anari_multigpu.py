"""GPU suite, ANARI boundary on a multi-GPU device: the device parameters "cudaDevices" / "multiGpuMode" put the
sort-last (fused slab frames) and sort-first paths behind the unchanged anariRenderFrame / anariMapFrame calls — one
process, peer access, z-slabs created by the field's finalize, compositing into the display GPU's frame
(the reference is single-GPU: VisRTXDevice.cpp:460-473, frame/Frame.cu:195-310).  Needs >= 2 GPUs."""
import numpy as np
import pytest

import dvr_harness as H
from test_gpu_anari import AnariScene, _errors
from visrtx_b200 import anari as A
from visrtx_b200 import scenes

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _frames(s, n):
    out = None
    for _ in range(n):
        s.render()
    color, w, h, _ = s.d.map_frame(s.frame, "channel.color")
    depth, _, _, _ = s.d.map_frame(s.frame, "channel.depth")
    obj, _, _, _ = s.d.map_frame(s.frame, "channel.objectId")
    return color, depth, obj


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sort_last_through_anari_matches_the_single_gpu_frame(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    vox = scenes.blobs_np(64)
    kw = dict(n=64, w=200, h=152, renderer="default", rate=0.5, vox=vox)
    one = AnariScene(**kw)
    many = AnariScene(gpus=list(range(world)), multi_gpu_mode="sortLast", **kw)
    for s in (one, many):
        s.d.set(s.volume, "unitDistance", A.FLOAT32, 0.5)
        s.d.commit(s.volume)
    c1, d1, o1 = _frames(one, 3)  # 3 accumulated frames
    cn, dn, on = _frames(many, 3)
    assert not _errors(many.d), many.d.messages
    assert many.d.get_property(many.d.handle, "cudaDeviceCount", A.INT32) == world  # (the device initialises lazily)
    assert many.d.get_property(many.frame, "numSamples", A.INT32) == 2
    d = np.abs(H.unpack_rgba8(cn) - H.unpack_rgba8(c1)).max(axis=-1)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 2, (float((d <= 1).mean()), int(d.max()))
    np.testing.assert_allclose(dn, d1, rtol=1e-6)
    assert np.array_equal(on, o1)
    # a moved camera restarts accumulation and changes the screen window of the volume
    for s in (one, many):
        pose = scenes.orbit_camera((-1,) * 3, (1,) * 3, 200, 152, az_deg=200.0, el_deg=-25.0, dist_scale=0.8)
        s.d.set(s.camera, "position", A.FLOAT32_VEC3, pose.position)
        s.d.set(s.camera, "direction", A.FLOAT32_VEC3, pose.direction)
        s.d.set(s.camera, "up", A.FLOAT32_VEC3, pose.up)
        s.d.commit(s.camera)
    c1, d1, _ = _frames(one, 1)
    cn, dn, _ = _frames(many, 1)
    d = np.abs(H.unpack_rgba8(cn) - H.unpack_rgba8(c1)).max(axis=-1)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 2
    np.testing.assert_allclose(dn, d1, rtol=1e-6)
    # a transfer-function edit re-finalises the per-GPU volumes
    for s in (one, many):
        s.d.set(s.volume, "unitDistance", A.FLOAT32, 0.2)
        s.d.commit(s.volume)
    c1, _, _ = _frames(one, 1)
    cn, _, _ = _frames(many, 1)
    d = np.abs(H.unpack_rgba8(cn) - H.unpack_rgba8(c1)).max(axis=-1)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 2
    assert not _errors(many.d), many.d.messages
    one.close()
    many.close()


def test_sort_first_through_anari_is_bit_identical():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(_gpus(), 4)
    kw = dict(n=48, w=160, h=120, renderer="default", rate=0.5)
    one = AnariScene(**kw)
    many = AnariScene(gpus=list(range(world)), multi_gpu_mode="sortFirst", **kw)
    c1, d1, o1 = _frames(one, 2)
    cn, dn, on = _frames(many, 2)
    assert not _errors(many.d), many.d.messages
    assert np.array_equal(cn, c1) and np.array_equal(dn, d1) and np.array_equal(on, o1)
    one.close()
    many.close()


def test_scenes_outside_the_distributed_paths_fall_back_to_the_display_gpu():
    """dpt renderer on a sort-last device: the frame is rendered on the display GPU from the whole field (uploaded on
    demand) — same image as a single-GPU device."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    kw = dict(n=40, w=96, h=96, renderer="dpt", rate=0.5, color_type=A.FLOAT32_VEC4)
    one = AnariScene(**kw)
    many = AnariScene(gpus=[0, 1], multi_gpu_mode="sortLast", **kw)
    c1, _, _ = _frames(one, 2)
    cn, _, _ = _frames(many, 2)
    assert np.array_equal(cn, c1)
    assert not _errors(many.d), many.d.messages
    one.close()
    many.close()
