"""Mixed scenes: surfaces in front of / behind volumes, lights, surface- and volume-attenuated shadow
rays, ambient occlusion — the CUDA path (dvr_render_scene) against O-gpu, i.e. the reference's own sampleLight /
computeAO / rayMarchVolume / accumResults headers driven by the restated surface branch of the raygen programs
(renderer/DirectLight_ptx.cu:64-218,294-418, Raycast_ptx.cu:60-179).

Tolerance: north_star's <= 2/255 per pixel, with one allowance the volume-only path does not need — OptiX's triangle
test is replaced by a double-precision test in the oracle and a fp32 Moeller-Trumbore in the product, so a pixel-sample
whose ray grazes a silhouette edge can hit in one and miss in the other.  The test therefore reports the outliers and
bounds their share (<= 0.3 % of the pixels) instead of demanding zero."""
import numpy as np
import pytest

import dvr_harness as H
import oracle_binding as ob
from visrtx_b200 import capi

pytestmark = pytest.mark.gpu

ZOO = H.mixed_scene_zoo()


def _report(name, got, want, scene):
    if scene.fmt == capi.DVR_FORMAT_FLOAT32_VEC4:
        d = np.abs(got["color"] - want["color"]).max(axis=-1) * 255.0
    else:
        d = np.abs(H.unpack_rgba8(got["color"]) - H.unpack_rgba8(want["color"])).max(axis=-1)
    frac = float((d <= 2).mean())
    print(f"{name}: max|d|={d.max():.1f}/255 frac(<=2/255)={frac:.5f} outliers={int((d > 2).sum())}")
    return d, frac


@pytest.mark.skipif(not ob.have_ref_gpu(), reason="oracle/_ref/libref_gpu_dvr.so not built")
@pytest.mark.parametrize("name", sorted(ZOO))
def test_mixed_scene_matches_reference_device_code(name):
    scene = ZOO[name]
    frames = 2 if "spp2" in name else 1
    got = H.render_cuda(scene, frames=frames)
    want = H.render_refgpu(scene, frames=frames)
    d, frac = _report(name, got, want, scene)
    assert frac >= 0.997, f"{name}: {int((d > 2).sum())} pixels differ by more than 2/255"
    assert float(np.median(d)) <= 1.0
    # depth / ids agree wherever the colour does (silhouette flips change them too)
    ok = d <= 2
    if "depth" in got:
        assert np.allclose(got["depth"][ok], want["depth"][ok], rtol=2e-4, atol=1e-4)
    for ch in ("primId", "objId", "instId"):
        if ch in got:
            assert (got[ch][ok] == want[ch][ok]).mean() >= 0.999, ch
    if "normal" in got:
        assert np.abs(got["normal"][ok] - want["normal"][ok]).max() <= 2e-3
        assert np.abs(got["albedo"][ok] - want["albedo"][ok]).max() <= 2.0 / 255.0 * 4


def test_pixels_without_a_surface_are_bit_identical_to_the_volume_frame():
    """A surface far outside the view changes nothing: the mixed-scene kernel runs the volume frame's arithmetic."""
    scene = H.default_scene(32, 96, 96, rate=0.5, field="blobs", integrator=capi.DVR_INTEGRATOR_DEFAULT,
                            num_iterations=2)
    plain = H.render_cuda(scene, frames=2)
    scene.surfaces = [{"geometry": "sphere", "vertex.position": [(50.0, 80.0, -120.0)], "radius": 0.5}]
    scene.lights = [{"type": "directional", "direction": (0, -1, 0)}]
    mixed = H.render_cuda(scene, frames=2)
    for k in plain:
        assert np.array_equal(plain[k], mixed[k]), k


def test_volume_casts_a_shadow_and_hides_behind_an_opaque_surface():
    """Known answers: (1) the floor under the volume is darker than the floor beside it (volumeAttenuation), (2) a
    pixel whose first hit is an opaque surface carries that surface's ids, and with no light and no ambient term it is
    black whatever volume lies in front (the reference's shading result replaces the volume segment's colour)."""
    import copy
    scene = copy.copy(ZOO["floor_balls_sun"])
    # the light travels to the right of the view (camera at (3.3, 2.4, 5.6) looking at the origin), so the volume's
    # shadow lies on floor the camera sees directly
    scene.surfaces = [dict(ZOO["floor_balls_sun"].surfaces[0],
                           **{"vertex.position": H._quad((-6, -1.3, -6), (-6, -1.3, 6), (6, -1.3, 6), (6, -1.3, -6))}),
                      ZOO["floor_balls_sun"].surfaces[1]]
    scene.lights = [{"type": "directional", "direction": (0.8, -0.6, -0.46), "irradiance": 3.0}]
    scene.ambient_samples = 0 # (the AO directions would differ between the two frames: the volume's shadow march draws)
    lit = H.render_cuda(scene)
    no_vol = copy.copy(scene)
    no_vol.volumes = []
    bare = H.render_cuda(no_vol)
    a = H.unpack_rgba8(lit["color"]).reshape(scene.height, scene.width, 4)[..., :3].sum(axis=-1)
    b = H.unpack_rgba8(bare["color"]).reshape(scene.height, scene.width, 4)[..., :3].sum(axis=-1)
    floor = (lit["objId"].reshape(scene.height, scene.width) == 11) & (bare["objId"].reshape(scene.height, scene.width) == 11)
    assert floor.sum() > 200
    darker = (a < b - 6) & floor
    assert darker.sum() > 50, "the volume casts no shadow on the floor"
    assert not ((a > b + 6) & floor).any()
    dark = copy.copy(scene)
    dark.lights, dark.surface_ambient_radiance = [], 0.0
    black = H.render_cuda(dark)
    hit = black["objId"] == 12
    assert hit.sum() > 20 and (H.unpack_rgba8(black["color"])[hit][:, :3] == 0).all()


def test_scene_argument_checks():
    scene = ZOO["floor_balls_sun"]
    cs = H.CudaScene(scene)
    try:
        p = H._params(scene, 0, -1)
        p.integrator = capi.DVR_INTEGRATOR_DPT
        sp, keep = scene.scene_params(cs.surfaces.handle)
        with pytest.raises(capi.DvrError) as e:
            capi.render_scene(p, scene.camera, cs.instances, cs.n, sp, cs.fb)
        assert e.value.code == capi.DVR_ERR_UNSUPPORTED
        n_prims, n_nodes = cs.surfaces.info()
        assert n_prims == 2 + 3 and n_nodes >= 1
    finally:
        cs.destroy()
    with pytest.raises(capi.DvrError) as e:
        capi.Surfaces.create([{"geometry": "triangle", "vertex.position": np.zeros((3, 3), np.float32),
                               "primitive.index": [(0, 1, 7)]}])
    assert e.value.code == capi.DVR_ERR_INVALID_ARGUMENT


# ---- the same through the ANARI C API -----------------------------------------------------------------------
def _anari_mixed(scene, renderer_subtype):
    """Builds `scene` (one volume, surfaces given as dicts, lights) through anari* calls as an application would and
    returns the mapped colour / depth / id channels of one frame."""
    from visrtx_b200 import anari as A
    d = A.Device()
    v = scene.volumes[0]
    n = v.voxels.shape[0]
    data = d.new_array3d(v.voxels, A.FLOAT32)
    field = d.new("SpatialField", "structuredRegular")
    d.set(field, "data", A.ARRAY3D, data)
    d.set(field, "origin", A.FLOAT32_VEC3, v.origin)
    d.set(field, "spacing", A.FLOAT32_VEC3, v.spacing)
    d.commit(field)
    volume = d.new("Volume", "transferFunction1D")
    from visrtx_b200 import scenes
    color = d.new_array1d(scenes.tsd_default_colormap(256), A.FLOAT32_VEC4)
    d.set(volume, "color", A.ARRAY1D, color)
    d.set(volume, "value", A.SPATIAL_FIELD, field)
    d.set(volume, "valueRange", A.FLOAT32_BOX1, v.value_range)
    d.set(volume, "unitDistance", A.FLOAT32, v.unit_distance)
    d.set(volume, "id", A.UINT32, v.vol_id)
    d.commit(volume)
    world = d.new("World")
    d.set(world, "volume", A.ARRAY1D, d.new_object_array([volume], A.VOLUME))
    zero_surfaces, instances = [], []
    for sdef in scene.surfaces:
        tri = sdef.get("geometry", "triangle") == "triangle"
        g = d.new("Geometry", "triangle" if tri else "sphere")
        pos = np.ascontiguousarray(sdef["vertex.position"], np.float32).reshape(-1, 3)
        d.set(g, "vertex.position", A.ARRAY1D, d.new_array1d(pos, A.FLOAT32_VEC3))
        if "primitive.index" in sdef:
            idx = np.ascontiguousarray(sdef["primitive.index"], np.uint32)
            d.set(g, "primitive.index", A.ARRAY1D, d.new_array1d(idx.reshape(-1, 3) if tri else idx.ravel(),
                                                                   A.UINT32_VEC3 if tri else A.UINT32))
        if "vertex.normal" in sdef:
            d.set(g, "vertex.normal", A.ARRAY1D, d.new_array1d(np.asarray(sdef["vertex.normal"], np.float32), A.FLOAT32_VEC3))
        if "vertex.radius" in sdef:
            d.set(g, "vertex.radius", A.ARRAY1D, d.new_array1d(np.asarray(sdef["vertex.radius"], np.float32), A.FLOAT32))
        if "radius" in sdef:
            d.set(g, "radius", A.FLOAT32, sdef["radius"])
        if "primitive.id" in sdef:
            d.set(g, "primitive.id", A.ARRAY1D, d.new_array1d(np.asarray(sdef["primitive.id"], np.uint32), A.UINT32))
        if sdef.get("cullBackfaces"):
            d.set(g, "cullBackfaces", A.BOOL, 1)
        d.commit(g)
        m = d.new("Material", "matte")
        col = sdef.get("color", (0.8, 0.8, 0.8))
        d.set(m, "color", A.FLOAT32_VEC4 if len(col) == 4 else A.FLOAT32_VEC3, col)
        d.set(m, "opacity", A.FLOAT32, sdef.get("opacity", 1.0))
        d.set(m, "alphaMode", A.STRING, sdef.get("alphaMode", "opaque"))
        d.set(m, "alphaCutoff", A.FLOAT32, sdef.get("alphaCutoff", 0.5))
        d.commit(m)
        s = d.new("Surface")
        d.set(s, "geometry", A.GEOMETRY, g)
        d.set(s, "material", A.MATERIAL, m)
        d.set(s, "id", A.UINT32, sdef.get("id", 0xFFFFFFFF))
        d.commit(s)
        if "transform" in sdef or sdef.get("instanceId", 0xFFFFFFFF) != 0xFFFFFFFF:
            grp = d.new("Group")
            d.set(grp, "surface", A.ARRAY1D, d.new_object_array([s], A.SURFACE))
            d.commit(grp)
            inst = d.new("Instance", "transform")
            d.set(inst, "group", A.GROUP, grp)
            d.set(inst, "id", A.UINT32, sdef.get("instanceId", 0xFFFFFFFF))
            rm = np.asarray(sdef.get("transform", capi.IDENTITY_3X4), np.float32).reshape(3, 4)
            d.set(inst, "transform", A.FLOAT32_MAT3x4, rm.T.copy().ravel())  # column-major 4x3
            d.commit(inst)
            instances.append(inst)
        else:
            zero_surfaces.append(s)
    if zero_surfaces:
        d.set(world, "surface", A.ARRAY1D, d.new_object_array(zero_surfaces, A.SURFACE))
    if instances:
        d.set(world, "instance", A.ARRAY1D, d.new_object_array(instances, A.INSTANCE))
    lights = []
    for l in scene.lights or []:
        lt = d.new("Light", l.get("type", "directional"))
        d.set(lt, "color", A.FLOAT32_VEC3, l.get("color", (1, 1, 1)))
        if l.get("type") == "point":
            d.set(lt, "position", A.FLOAT32_VEC3, l["position"])
            d.set(lt, "intensity", A.FLOAT32, l.get("intensity", 1.0))
        else:
            d.set(lt, "direction", A.FLOAT32_VEC3, l["direction"])
            d.set(lt, "irradiance", A.FLOAT32, l.get("irradiance", 1.0))
        d.commit(lt)
        lights.append(lt)
    if lights:
        d.set(world, "light", A.ARRAY1D, d.new_object_array(lights, A.LIGHT))
    d.commit(world)
    cam = d.new("Camera", "perspective")
    lo, hi = v.bounds()
    pose = scenes.orbit_camera(lo, hi, scene.width, scene.height)
    for k, val in (("position", pose.position), ("direction", pose.direction), ("up", pose.up)):
        d.set(cam, k, A.FLOAT32_VEC3, val)
    d.set(cam, "fovy", A.FLOAT32, pose.fovy)
    d.set(cam, "aspect", A.FLOAT32, pose.aspect)
    d.commit(cam)
    r = d.new("Renderer", renderer_subtype)
    d.set(r, "background", A.FLOAT32_VEC4, scene.background)
    d.set(r, "volumeSamplingRate", A.FLOAT32, scene.volume_sampling_rate)
    d.set(r, "pixelSamples", A.INT32, scene.num_iterations)
    d.set(r, "ambientColor", A.FLOAT32_VEC3, scene.ambient_color)
    d.set(r, "ambientRadiance", A.FLOAT32, scene.surface_ambient_radiance)
    d.set(r, "ambientSamples", A.INT32, scene.ambient_samples)
    d.set(r, "cullTriangleBackfaces", A.BOOL, 1 if scene.cull_triangle_backfaces else 0)
    d.commit(r)
    f = d.new("Frame")
    d.set(f, "size", A.UINT32_VEC2, (scene.width, scene.height))
    d.set(f, "channel.color", A.DATA_TYPE, A.UFIXED8_RGBA_SRGB)
    d.set(f, "channel.depth", A.DATA_TYPE, A.FLOAT32)
    for ch in ("objectId", "instanceId", "primitiveId"):
        d.set(f, "channel." + ch, A.DATA_TYPE, A.UINT32)
    d.set(f, "renderer", A.RENDERER, r)
    d.set(f, "camera", A.CAMERA, cam)
    d.set(f, "world", A.WORLD, world)
    d.commit(f)
    d.render(f)
    d.wait(f)
    out = {}
    for ch, key in (("color", "color"), ("depth", "depth"), ("objectId", "objId"), ("instanceId", "instId"),
                    ("primitiveId", "primId")):
        a, w, h, _ = d.map_frame(f, "channel." + ch)
        out[key] = np.array(a, copy=True)
    bounds = d.get_property(world, "bounds", A.FLOAT32_BOX3)
    errors = [m for m in d.messages if m[0] <= A.SEVERITY_ERROR]
    d.close()
    return out, bounds, errors


@pytest.mark.parametrize("name,renderer", [("floor_balls_sun", "default"), ("raycast_mesh_instance", "raycast"),
                                           ("translucent_sheet", "directLight")])
def test_mixed_scene_through_anari_equals_cabi(name, renderer):
    """World 'surface' / 'light' / instanced groups, Geometry(triangle, sphere), Material(matte), Surface, Light objects
    of the ANARI device feed the same launch: the frame equals the C-ABI frame bit for bit."""
    scene = ZOO[name]
    got, bounds, errors = _anari_mixed(scene, renderer)
    assert not errors, errors
    want = H.render_cuda(scene)
    # (the device normalises the light direction once more: an ulp there may move an 8-bit colour by one step)
    d = np.abs(H.unpack_rgba8(got["color"]) - H.unpack_rgba8(want["color"])).max(axis=-1)
    assert d.max() <= 1 and (d == 0).mean() >= 0.999, (d.max(), (d == 0).mean())
    want["instId"] = np.where(want["instId"] == 3, 0xFFFFFFFF, want["instId"]).astype(np.uint32)  # world-level volume
    for k in ("depth", "objId", "instId", "primId"):
        assert np.array_equal(got[k], want[k]), k
    b = np.asarray(bounds, np.float32)
    assert b[0] <= -3.0 and b[3] >= 3.0 and b[1] <= -1.3 + 1e-6  # the floor quad extends the world bounds
