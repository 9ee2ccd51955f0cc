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
