"""GPU suite: edge cases of the marching path against O-gpu (the reference's device code, live) and O-cpu —
NaN voxels, degenerate value ranges, extreme sampling rates, axis-aligned rays, 1-pixel frames, flat volumes,
rays starting inside the volume, fully transparent and fully opaque transfer functions.  Same tolerances as the
parity suite; skipping on / off stays bit-identical on every one of them."""
import numpy as np
import pytest

import dvr_harness as H
import oracle_binding as ob
from visrtx_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def _scene(vox, w=64, h=48, rate=0.5, tf=None, value_range=(0.0, 1.0), integrator=capi.DVR_INTEGRATOR_DEFAULT, cam=None,
           spacing=None, unit_distance=None, **kw):
    nz, ny, nx = vox.shape
    sp = spacing or (2.0 / max(nx - 1, 1),) * 3
    tf = tf if tf is not None else capi.tf_discretize(color=scenes.tsd_default_colormap(256), value_range=value_range)
    v = H.VolumeDesc(vox, origin=(-1.0, -1.0, -1.0), spacing=sp, tf=tf, value_range=value_range,
                     unit_distance=unit_distance or sp[0], vol_id=5, inst_id=6)
    lo, hi = v.bounds()
    if cam is None:
        pose = scenes.orbit_camera(lo, hi, w, h, dist_scale=1.1)
        cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    return H.SceneDesc([v], w, h, cam, volume_sampling_rate=rate, integrator=integrator, **kw)


def _compare(scene, frames=2, name="", oracle_cpu=True):
    got = H.render_cuda(scene, frames=frames)
    skip = H.render_cuda(scene, frames=frames, skip=True)
    for k in got:
        assert np.array_equal(got[k], skip[k], equal_nan=True), (name, k)
    refs = []
    if ob.have_ref_gpu():
        refs.append(("O-gpu", H.render_refgpu(scene, frames=frames)))
    if oracle_cpu:
        refs.append(("O-cpu", H.render_oracle(scene, frames=frames)))
    for label, want in refs:
        d = np.abs(H.unpack_rgba8(got["color"]) - H.unpack_rgba8(want["color"])).max(axis=-1)
        assert (d <= 2).mean() >= 0.999 and d.max() <= 6, (name, label, float((d <= 2).mean()), int(d.max()))
        np.testing.assert_allclose(got["depth"], want["depth"], rtol=2e-5, atol=1e-5, err_msg=f"{name} {label}")
        assert (got["objId"] == want["objId"]).mean() >= 0.999, (name, label)
    return got


def test_nan_voxels_are_skipped_like_the_reference():
    vox = scenes.marschner_lobb_np(32)
    vox[10:20, 8:16, 12:24] = np.nan
    vox[0, 0, 0] = np.nan
    got = _compare(_scene(vox), name="nan block")
    assert np.isfinite(got["accum"]).all()
    allnan = np.full((8, 8, 8), np.nan, np.float32)
    got = _compare(_scene(allnan), name="all NaN")
    bg = H.render_cuda(_scene(np.zeros((8, 8, 8), np.float32), tf=np.zeros((256, 4), np.float32)))
    assert np.array_equal(got["color"], H.render_cuda(_scene(allnan), frames=2)["color"])
    assert len(np.unique(got["color"])) == 1 and np.unique(got["color"])[0] == np.unique(bg["color"])[0]


def test_degenerate_and_inverted_value_ranges():
    vox = scenes.blobs_np(24)
    for vr in ((0.5, 0.5), (1.0, 0.0), (-1e30, 1e30), (0.25, 0.25000003)):
        tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256))
        _compare(_scene(vox, tf=tf, value_range=vr), name=f"valueRange {vr}")


@pytest.mark.parametrize("rate", [1e-3, 0.01, 10.0])
def test_extreme_sampling_rates(rate):
    # 1e-3: one lattice point per ray at most; 10: twenty samples per voxel (Renderer.cpp:166-167 clamps to these)
    _compare(_scene(scenes.marschner_lobb_np(24), w=48, h=32, rate=rate, unit_distance=0.5), name=f"rate {rate}")


def test_axis_aligned_rays_and_orthographic_grazing():
    vox = scenes.blobs_np(20)
    for direction, up in (((0, 0, -1), (0, 1, 0)), ((1, 0, 0), (0, 1, 0)), ((0, -1, 0), (0, 0, 1))):
        pos = tuple(-3.0 * c for c in direction)
        cam = capi.camera_orthographic(pos, direction, up, 2.5, 64 / 48)
        _compare(_scene(vox, cam=cam), name=f"ortho {direction}")
    # perspective looking exactly down -z from the axis: the centre pixel column has dir.x == 0 in raycast mode
    cam = capi.camera_perspective((0.0, 0.0, 4.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 0.8, 64 / 48)
    _compare(_scene(vox, cam=cam, integrator=capi.DVR_INTEGRATOR_RAYCAST), frames=1, name="perspective on axis")


def test_tiny_frames_and_flat_volumes():
    vox = scenes.marschner_lobb_np(16)
    for w, h in ((1, 1), (1, 7), (9, 1), (3, 2)):
        _compare(_scene(vox, w=w, h=h), name=f"{w}x{h}")
    flat = np.ascontiguousarray(scenes.marschner_lobb_np(16)[:1])  # one slice: zero-thickness bounds, never hit
    got = _compare(_scene(flat, spacing=(2.0 / 15,) * 3), name="one slice")
    assert len(np.unique(got["color"])) == 1
    two = np.ascontiguousarray(scenes.marschner_lobb_np(16)[:2])
    _compare(_scene(two, spacing=(2.0 / 15,) * 3), name="two slices")


def test_camera_inside_looking_out_and_behind():
    vox = scenes.blobs_np(24)
    inside = capi.camera_perspective((0.1, -0.2, 0.3), (0.3, 0.2, -1.0), (0.0, 1.0, 0.0), 1.2, 64 / 48)
    _compare(_scene(vox, cam=inside), name="inside")
    away = capi.camera_perspective((0.0, 0.0, 4.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.8, 64 / 48)
    got = _compare(_scene(vox, cam=away), name="looking away")
    assert len(np.unique(got["color"])) == 1  # background only


def test_transparent_and_opaque_transfer_functions():
    vox = scenes.marschner_lobb_np(24)
    clear = np.zeros((256, 4), np.float32)
    clear[:, :3] = 0.7
    got = _compare(_scene(vox, tf=clear), name="alpha 0")
    assert len(np.unique(got["color"])) == 1
    solid = np.ones((256, 4), np.float32)
    _compare(_scene(vox, tf=solid, unit_distance=1e-3), name="alpha 1, tiny unitDistance")  # ERT after one sample
    _compare(_scene(vox, tf=solid, unit_distance=1e6), name="alpha 1, huge unitDistance")  # pow(0, ~0)


def test_the_test_renderer_paints_ray_directions():
    """`test` renderer (renderer/Test_ptx.cu): colour = jittered primary ray direction, depth 1, no scene access."""
    vox = scenes.blobs_np(16)
    s = _scene(vox, integrator=capi.DVR_INTEGRATOR_TEST, fmt=capi.DVR_FORMAT_FLOAT32_VEC4,
               channels=("depth", "objId", "albedo", "normal"))
    got = H.render_cuda(s, frames=2)
    want = H.render_oracle(s, frames=2)
    np.testing.assert_allclose(got["color"], want["color"], atol=2e-6)
    np.testing.assert_allclose(got["normal"], want["normal"], atol=4e-6)
    assert np.all(got["depth"] == 1.0) and np.all(got["objId"] == 0xFFFFFFFF)
    acc = got["accum"][:, :3].reshape(s.height, s.width, 3)[s.height // 2, s.width // 2] / 2  # two frames
    d = np.asarray(s.camera.dir)
    assert np.abs(acc - d / (1 + max(0.0, d.max()))).max() < 0.03  # the tonemapped direction of the centre ray
    if ob.have_ref_gpu():
        ref = H.render_refgpu(s, frames=2)
        for k in ("depth", "objId"):
            assert np.array_equal(got[k], ref[k]), k
        for k in ("color", "accum", "albedo", "normal"):  # same ray set-up as every other renderer: last-bit agreement
            np.testing.assert_allclose(got[k], ref[k], atol=1e-6, err_msg=k)
            assert (got[k] == ref[k]).mean() > 0.9, k
    # through ANARI
    from test_gpu_anari import AnariScene
    from visrtx_b200 import anari as A
    a = AnariScene(16, 64, 48, "test", 0.5, color_type=A.FLOAT32_VEC4, channels=("depth",))
    a.render()
    color, w, h, _ = a.d.map_frame(a.frame, "channel.color")
    assert not [m for m in a.d.messages if m[0] <= A.SEVERITY_WARNING], a.d.messages
    assert np.all(np.asarray(color).reshape(-1, 4)[:, 3] == 1.0)
    a.close()


def test_in_place_refresh_refuses_what_it_cannot_keep():
    """dvr_field_update_structured only keeps the device objects of a WHOLE structuredRegular field of unchanged
    element type: slabs, NanoVDB grids, FLOAT64 and type changes answer DVR_ERR_UNSUPPORTED and leave the field
    usable (the caller destroys and creates instead, as the ANARI device does)."""
    from visrtx_b200 import nvdb_writer
    n = 24
    vox = scenes.marschner_lobb_np(n)
    resident = np.ascontiguousarray(vox[7:17])
    slab = capi.Field.create_slab(resident.ctypes.data, False, capi.DVR_FLOAT32, (n, n, n), 8, 16, (0, 0, 0), (1, 1, 1))
    grid = nvdb_writer.fog_sphere(radius=10.0, voxel_size=1.0, half_width=3.0)
    nv = capi.Field.create_nanovdb(grid.ctypes.data, grid.nbytes)
    f64 = vox.astype(np.float64)
    dbl = capi.Field.create_structured(f64.ctypes.data, False, capi.DVR_FLOAT64, (n, n, n), (0, 0, 0), (1, 1, 1))
    whole = capi.Field.create_structured(vox.ctypes.data, False, capi.DVR_FLOAT32, (n, n, n), (0, 0, 0), (1, 1, 1))
    try:
        for fld, dt, data in ((slab, capi.DVR_FLOAT32, vox), (nv, capi.DVR_FLOAT32, vox), (dbl, capi.DVR_FLOAT64, f64),
                              (whole, capi.DVR_UFIXED8, vox)):
            with pytest.raises(capi.DvrError) as e:
                fld.update_structured(data.ctypes.data, False, dt, (0, 0, 0), (1, 1, 1))
            assert e.value.code == capi.DVR_ERR_UNSUPPORTED
        before = whole.value_range()
        half = np.ascontiguousarray(vox * np.float32(0.5))  # kept alive across the call
        whole.update_structured(half.ctypes.data, False, capi.DVR_FLOAT32, (0, 0, 0), (1, 1, 1))
        after = whole.value_range()
        assert after == (np.float32(before[0]) * np.float32(0.5), np.float32(before[1]) * np.float32(0.5))
    finally:
        for fld in (slab, nv, dbl, whole):
            fld.destroy()
