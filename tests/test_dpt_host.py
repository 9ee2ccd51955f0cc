"""CPU suite: the `dpt` renderer's delta (Woodcock) tracking — O-cpu pinned by known answers.

The reference's dpt raygen (renderer/DiffusePathTracer_ptx.cu:82-215) is stochastic and its free-path sampler
(gpu/volumeIntegration.h:167-238) has a closed-form expectation in a homogeneous medium, which gives
known-answer tests that need neither a GPU nor the reference binary:
  * P(no collision along a chord of length L) = exp(-alpha * L / stepSize)   (stepSize = min(spacing)/2)
  * with a white transfer function every scattered path carries Lw == 1, so a pixel is either the background
    or exactly the ambient radiance
  * the majorant grid must bound the classified opacity of every sample inside its cell (else the tracker is
    biased) — checked against O-cpu's own texture-unit model at random points.
O-cpu is test infrastructure; the product's CUDA tracker is compared with it (and with O-gpu) in
tests/test_gpu_dpt.py.
"""
import os

import numpy as np
import pytest

import dvr_harness as H
import oracle_binding as ob
from visrtx_b200 import capi, scenes


def _const_scene(alpha, color, n=17, wh=32, bg=(1.0, 1.0, 1.0, 1.0), **kw):
    vox = np.full((n, n, n), 0.5, np.float32)
    tf = np.zeros((256, 4), np.float32)
    tf[:, :3] = color
    tf[:, 3] = alpha
    v = H.VolumeDesc(vox, origin=(0.0, 0.0, 0.0), spacing=(1.0, 1.0, 1.0), tf=tf, unit_distance=1.0)
    lo, hi = v.bounds()
    pose = scenes.orbit_camera(lo, hi, wh, wh, dist_scale=0.9)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    return H.SceneDesc([v], wh, wh, cam, fmt=capi.DVR_FORMAT_FLOAT32_VEC4, integrator=capi.DVR_INTEGRATOR_DPT,
                       background=bg, channels=("depth", "objId"), **kw)


def _chords(scene):
    """Chord length of the pixel-centre ray through volume 0's box, per pixel (row-major)."""
    c = scene.camera
    w, h = scene.width, scene.height
    sx = (np.arange(w) + 0.5) / w
    sy = (np.arange(h) + 0.5) / h
    SX, SY = np.meshgrid(sx, sy)
    d = (np.asarray(c.p00)[None, None] + SX[..., None] * np.asarray(c.du) + SY[..., None] * np.asarray(c.dv))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.asarray(c.pos, np.float64)
    lo, hi = scene.volumes[0].bounds()
    with np.errstate(divide="ignore"):
        t0 = (lo - o) / d
        t1 = (hi - o) / d
    tn = np.minimum(t0, t1).max(-1)
    tf = np.maximum(t0, t1).min(-1)
    return np.maximum(tf - np.maximum(tn, 0.0), 0.0).ravel()


def test_dpt_transmittance_of_a_homogeneous_medium():
    alpha, frames = 0.04, 96
    s = _const_scene(alpha, (0.0, 0.0, 0.0))
    out = H.render_oracle(s, frames=frames)
    # black albedo: a collided path contributes 0, a free path the white background, tonemapped 1/(1+1) = .5
    t_hat = out["accum"][:, 0] / frames * 2.0
    L = _chords(s)
    t_ref = np.exp(-alpha * L / 0.5)
    inside = L > 1.0  # away from the silhouette, where the jittered ray's chord differs from the centre ray's
    assert inside.sum() > 200
    sigma = np.sqrt(np.maximum(t_ref * (1 - t_ref), 1e-4) / frames)
    z = np.abs(t_hat - t_ref)[inside] / sigma[inside]
    assert np.percentile(z, 99) < 4.0, np.percentile(z, 99)
    assert abs(t_hat[inside].mean() - t_ref[inside].mean()) < 0.01
    # outside the box: exactly the background (silhouette pixels excepted: the jittered ray may clip a corner)
    assert np.mean(t_hat[L == 0.0] == 1.0) > 0.9


def test_dpt_white_albedo_carries_unit_throughput():
    s = _const_scene(0.3, (1.0, 1.0, 1.0), bg=(0.25, 0.5, 0.75, 1.0), max_depth=256, ambient_radiance=0.6)
    out = H.render_oracle(s, frames=1)
    rgb = out["color"][:, :3]
    hit = np.all(np.abs(rgb - 0.6) < 1e-5, axis=1)
    bg = np.all(np.abs(rgb - np.array([0.25, 0.5, 0.75], np.float32)) < 1e-5, axis=1)
    assert np.all(hit | bg)
    assert hit.sum() > 300 and bg.sum() > 100
    assert np.all(out["color"][:, 3] == 1.0)
    # the reference's raygen never records depth / ids for this renderer (its `depth == 0` test runs after the
    # increment): depth stays at its reset value, ids at the reset 0
    assert np.all(out["depth"] == np.finfo(np.float32).max)
    assert np.all(out["objId"] == 0)


def test_dpt_max_depth_terminates_paths():
    # dense white medium, maxDepth 1: a path that scatters twice is killed (Lw = 0)
    s1 = _const_scene(0.9, (1.0, 1.0, 1.0), bg=(0.0, 0.0, 0.0, 1.0), max_depth=1)
    s2 = _const_scene(0.9, (1.0, 1.0, 1.0), bg=(0.0, 0.0, 0.0, 1.0), max_depth=64)
    a = H.render_oracle(s1, frames=1)["color"][:, 0]
    b = H.render_oracle(s2, frames=1)["color"][:, 0]
    assert a.sum() < 0.6 * b.sum()
    assert set(np.unique(np.round(a, 5))) <= {0.0, 1.0}


def test_dpt_path_state_persists_across_pixel_samples():
    """PathData is declared outside the numIterations loop (DiffusePathTracer_ptx.cu:87): once a sample
    collided, depth stays > 0 and Lw keeps its value for the remaining samples of that launch."""
    s = _const_scene(0.5, (0.0, 0.0, 0.0), num_iterations=8)
    out = H.render_oracle(s, frames=1)
    L = _chords(s)
    # black albedo: after the first collision every later sample returns Lw * ambient = 0, never the background
    thick = L > 8.0
    assert out["accum"][thick, 0].max() == 0.0


def test_dda_grid_geometry_and_conservative_majorants():
    s = H.default_scene(40, 32, 32, integrator=capi.DVR_INTEGRATOR_DPT)
    (dims, maj), = H.oracle_dda_grids(s)
    assert dims == (3, 3, 3)  # ceil(40/16), UniformGrid.cu:152-154
    v = s.volumes[0]
    lo, hi = v.bounds()
    rng = np.random.default_rng(5)
    p = (lo + (hi - lo) * rng.random((20000, 3))).astype(np.float32)
    sp = np.asarray(v.spacing, np.float32)
    tc = ((p - np.asarray(v.origin, np.float32)) + 0.5 * sp) / (sp * np.asarray(v.dims, np.float32))
    val = ob.tex3d(v.voxels, tc[:, 0], tc[:, 1], tc[:, 2])
    a = ob.tex1d_tf(v.tf, np.clip(val, 0, 1).astype(np.float32))[:, 3]
    cell = np.minimum(((p - lo) / (hi - lo) * np.asarray(dims)).astype(int), np.asarray(dims) - 1)
    m = maj[cell[:, 2], cell[:, 1], cell[:, 0]]
    assert np.all(a <= m + 1e-7)
    assert maj.max() <= v.tf[:, 3].max() + 1e-7


def test_dda_grid_nanovdb_geometry():
    from visrtx_b200 import nvdb_writer
    blob = nvdb_writer.fog_sphere(50.0)
    tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256), opacity=np.linspace(0, 1, 16, dtype=np.float32))
    v = H.VolumeDesc(np.zeros((1, 1, 1), np.float32), tf=tf, nvdb=blob)
    lo, hi = v.bounds()
    pose = scenes.orbit_camera(lo, hi, 16, 16)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    s = H.SceneDesc([v], 16, 16, cam, integrator=capi.DVR_INTEGRATOR_DPT)
    (dims, maj), = H.oracle_dda_grids(s)
    n = int(round(float(hi[0] - lo[0])))
    assert dims == ((n + 15) // 16,) * 3  # NvdbRegularField.cpp:153: the index bounding box's extent
    assert maj.min() < 1e-3 and maj.max() > 0.9  # corner cells of a sphere's bounding box are empty


@pytest.mark.parametrize("kind", H.DPT_KINDS)
def test_oracle_dpt_matches_golden_reference_tracker(kind):
    """tests/golden/refgpu_dpt.npz was rendered ON A B200 by O-gpu — the reference's unmodified
    sampleDistanceAllVolumes / _sampleDistance / dda3 / sampleUnitSphere / accumResults — over this grid.
    Same Philox stream and same statement, so O-cpu reproduces it pixel for pixel except where an ulp of
    libm-vs-CUDA logf/sinf/cosf or of the software texture unit flips an accept/reject test."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgpu_dpt.npz"))
    s = H.dpt_scene(kind)
    one = H.render_oracle(s, frames=1)
    acc = H.render_oracle(s, frames=H.DPT_GOLDEN_FRAMES)
    assert H.frac_close(one["color"], g[f"{kind}/color1"], s.fmt) >= 0.99, kind
    assert float(np.mean(one["color"] == g[f"{kind}/color1"])) >= 0.95, kind
    mae = np.abs(acc["accum"] - g[f"{kind}/accum"]).mean() / H.DPT_GOLDEN_FRAMES
    assert mae < 1e-3, (kind, mae)
