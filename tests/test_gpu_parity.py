"""GPU suite: the CUDA path, called through the C-ABI, against
  * the committed O-gpu golden renders (reference device code on a B200),
  * O-gpu run live on this GPU (when oracle/_ref/libref_gpu_dvr.so travelled with the snapshot),
  * O-cpu on the same seeded inputs,
and the size-independent properties of the path at BASELINE sizes.
Tolerances: north_star — per pixel <= 2/255 after tonemap, PSNR >= 45 dB (ERT-threshold outliers as in
test_oracle_golden.py); the macrocell-skipping and sort-first variants must be BIT-identical."""
import os

import numpy as np
import pytest

import dvr_harness as H
import oracle_binding as ob
from visrtx_b200 import capi, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ZOO = H.scene_zoo()


def _pix_diff(a, b, fmt):
    if fmt == capi.DVR_FORMAT_FLOAT32_VEC4:
        return np.abs(a - b).max(axis=-1) * 255.0
    return np.abs(H.unpack_rgba8(a) - H.unpack_rgba8(b)).max(axis=-1)


REPORT = {}  # per-scene figures, written to gpurun_out/parity_report.json at the end of the module


def _check(got, want, scene, frames=1, strict=True, name="", against=""):
    """north_star tolerance: every pixel <= 2/255 after tonemap, PSNR >= 45 dB.  strict=False (comparisons with the
    CPU restatement only, never with reference renders): up to 0.1 % of the pixels may differ by <= 6/255 — rays whose
    early termination at opacity 0.99 flips by one sample between the software and the hardware texture filter; the
    count is recorded per scene in the parity report."""
    max_d, psnr = H.compare_color(got["color"], want["color"], scene.fmt)
    d = _pix_diff(got["color"], want["color"], scene.fmt)
    REPORT[f"{against}:{name}"] = {"max_abs_255": float(max_d), "psnr_db": float(psnr), "pixels": int(d.size),
                                   "pixels_gt_2": int((d > 2).sum()), "pixels_identical": int((d == 0).sum())}
    assert psnr >= 45.0, (name, psnr)
    assert (d <= 2).mean() >= 0.999, (name, float((d <= 2).mean()))
    assert max_d <= (2.0 if strict else 6.0), (name, max_d)
    if "depth" in got and "depth" in want:
        np.testing.assert_allclose(got["depth"], want["depth"], rtol=2e-5, atol=1e-5)
    for key in ("primId", "objId", "instId"):
        if key in got and key in want:
            assert (got[key] == want[key]).mean() >= 0.999, (name, key)
    for key in ("albedo", "normal"):
        if key in got and key in want:
            assert np.percentile(np.abs(got[key] - want[key]), 99.9) <= 0.02 * frames * scene.num_iterations


@pytest.mark.parametrize("name", sorted(ZOO))
def test_cuda_matches_golden_reference_renders(name):
    scene, frames, cb = ZOO[name]
    g = np.load(os.path.join(GOLD, "refgpu_scenes.npz"))
    want = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    got = H.render_cuda(scene, frames=frames, checkerboard=cb)
    _check(got, want, scene, frames, name=name, against="golden-O-gpu")


@pytest.mark.parametrize("name", sorted(ZOO))
def test_cuda_matches_oracle_cpu(name):
    scene, frames, cb = ZOO[name]
    got = H.render_cuda(scene, frames=frames, checkerboard=cb)
    want = H.render_oracle(scene, frames=frames, checkerboard=cb)
    _check(got, want, scene, frames, strict=False, name=name, against="O-cpu")


def teardown_module(module):
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.mark.parametrize("name", sorted(ZOO))
def test_cuda_matches_reference_device_code_live(name):
    if not ob.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu_dvr.so not present")
    scene, frames, cb = ZOO[name]
    got = H.render_cuda(scene, frames=frames, checkerboard=cb)
    want = H.render_refgpu(scene, frames=frames, checkerboard=cb)
    _check(got, want, scene, frames, strict=True, name=name, against="live-O-gpu")
    # most pixels agree to the last bit of the float accumulation buffer
    # (the thin-lens + instance-transform and the NanoVDB scenes have more places where FMA contraction may differ)
    bit_identical = float((got["accum"] == want["accum"]).all(axis=-1).mean())
    REPORT[f"live-O-gpu:{name}"]["accum_bit_identical_frac"] = bit_identical
    # Scenes whose camera arithmetic has more than one legal FMA contraction (image-region mix, orthographic origin,
    # thin lens, instance transforms) and the NanoVDB scenes agree in fewer last bits; their 8-bit images are held to
    # the same <= 2/255 above.  measured: camera_inside_region 0.24, noise_ortho_* 0.71 (parity report)
    loose = name in ("two_volumes_lens", "camera_inside_region", "noise_ortho_linear", "noise_ortho_nearest") \
        or name.startswith("nvdb")
    assert bit_identical > (0.2 if loose else 0.9), (name, bit_identical)


def test_config_c1_full_size():
    """BASELINE config 1: 64^3 Marschner-Lobb, 512x512, 1 spp, raycast."""
    scene = H.default_scene(64, 512, 512, rate=0.5)
    got = H.render_cuda(scene)
    want = H.render_oracle(scene)
    _check(got, want, scene, strict=False, name="C1", against="O-cpu")
    if ob.have_ref_gpu():
        _check(got, H.render_refgpu(scene), scene, strict=True, name="C1", against="live-O-gpu")


@pytest.mark.parametrize("name", ["ml48_raycast_r1.0", "blobs48_translucent", "two_volumes_lens", "blobs32_u8",
                                  "nvdb_fog_r20", "nvdb_fog_r12_vs025", "nvdb_fp4_r14", "nvdb_fpn_r14"])
def test_macrocell_skipping_is_bit_identical(name):
    scene, frames, cb = ZOO[name]
    if name.startswith("blobs"):
        for v in scene.volumes:  # make most of the volume fully transparent
            v.tf = capi.tf_discretize(color=scenes.sparse_colormap(256, 0.4))
    a = H.render_cuda(scene, frames=frames, checkerboard=cb, skip=False)
    b = H.render_cuda(scene, frames=frames, checkerboard=cb, skip=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), (name, k)


def test_macrocell_skipping_actually_skips_and_majorants_are_conservative():
    import torch
    n = 96
    vox = scenes.shells_np(n, n_shells=4, radius_frac=0.12, width_frac=0.01)
    sp = 2.0 / (n - 1)
    tf = capi.tf_discretize(color=scenes.sparse_colormap(256, 0.3))
    v = H.VolumeDesc(vox, origin=(-1, -1, -1), spacing=(sp,) * 3, tf=tf, unit_distance=4 * sp)
    pose = scenes.orbit_camera((-1, -1, -1), (1, 1, 1), 256, 256, dist_scale=1.0)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    scene = H.SceneDesc([v], 256, 256, cam, volume_sampling_rate=1.0)
    cs = H.CudaScene(scene)
    try:
        s0 = cs.render(stats=True, skip=False)
        a = cs.download()
        s1 = cs.render(stats=True, skip=True)
        b = cs.download()
        # fetches are issued in batches, so up to DVR_BATCH-1 prefetched samples past early termination
        # are counted per ray; apart from that the lattice is identical
        assert abs(s1["samplesTaken"] + s1["samplesSkipped"] - s0["samplesTaken"]) <= 4 * s0["raysHit"]
        assert s1["samplesSkipped"] > 0.5 * s0["samplesTaken"]
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        # macrocell ranges: conservative w.r.t. a numpy min/max over the cell + apron
        (gx, gy, gz), ptr = cs.fields[0].macrocells()
        assert (gx, gy, gz) == (6, 6, 6)
        rng = torch.empty((gz, gy, gx, 2), dtype=torch.float32, device="cuda")
        import ctypes
        torch.cuda.synchronize()
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(rng.data_ptr()), ctypes.c_void_p(ptr),
                                               ctypes.c_size_t(rng.numel() * 4), ctypes.c_int(3))
        r = rng.cpu().numpy()
        for cz in range(gz):
            for cy in range(gy):
                for cx in range(gx):
                    blk = vox[max(cz * 16 - 1, 0):cz * 16 + 18, max(cy * 16 - 1, 0):cy * 16 + 18,
                              max(cx * 16 - 1, 0):cx * 16 + 18]
                    assert r[cz, cy, cx, 0] == blk.min() and r[cz, cy, cx, 1] == blk.max()
        lo, hi = cs.fields[0].value_range()
        assert lo == vox.min() and hi == vox.max()
    finally:
        cs.destroy()


def test_sort_first_tiles_reassemble_bit_identically():
    """Rendering tile rows rank-by-rank into one buffer gives exactly the single-GPU frame."""
    scene, _, _ = ZOO["ml48_raycast_r0.5"]
    full = H.render_cuda(scene)
    for ranks in (2, 3, 8):
        cs = H.CudaScene(scene)
        try:
            for r in range(ranks):
                cs.render(tile_rank=r, tile_ranks=ranks)
            part = cs.download()
        finally:
            cs.destroy()
        for k in full:
            assert np.array_equal(full[k], part[k]), (ranks, k)


@pytest.mark.parametrize("nslabs", [2, 4])
def test_sort_last_slabs_composite_to_the_single_gpu_image(nslabs):
    """z-slabs rendered on the global lattice + `over` compositing + resolve == the one-pass frame."""
    import torch
    scene = H.default_scene(64, 160, 120, rate=0.5, field="blobs")
    scene.volumes[0].unit_distance = 0.5
    want = H.render_cuda(scene)
    n = scene.width * scene.height
    nz = 64
    bounds = [round(i * nz / nslabs) for i in range(nslabs + 1)]
    parts = []
    p = H._params(scene, 0, -1)
    for i in range(nslabs):
        cs = H.CudaScene(scene, slab=(bounds[i], bounds[i + 1]))
        rgba = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        depth = torch.zeros(n, dtype=torch.float32, device="cuda")
        capi.render_partial(p, scene.camera, cs.instances, rgba.data_ptr(), depth.data_ptr())
        torch.cuda.synchronize()
        parts.append((rgba, depth, cs))
    # view order along z: the camera of default_scene looks from +z towards -z => slab nslabs-1 is in front
    cam_z_positive = scene.camera.pos[2] > 0
    order = list(range(nslabs))[::-1] if cam_z_positive else list(range(nslabs))
    front_rgba, front_depth, _ = parts[order[0]]
    for j in order[1:]:
        capi.composite_over(front_rgba.data_ptr(), front_depth.data_ptr(), parts[j][0].data_ptr(),
                            parts[j][1].data_ptr(), 0, n, False)
    out = H.CudaScene(scene)
    try:
        capi.resolve(p, front_rgba.data_ptr(), front_depth.data_ptr(), scene.volumes[0].vol_id,
                     scene.volumes[0].inst_id, out.fb, 0, n)
        got = out.download()
    finally:
        out.destroy()
        for _, _, cs in parts:
            cs.destroy()
    d = _pix_diff(got["color"], want["color"], scene.fmt)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 3
    np.testing.assert_allclose(got["depth"], want["depth"], rtol=1e-6)
    assert np.array_equal(got["objId"], want["objId"])


def test_progressive_accumulation_converges_and_counts():
    """64 accumulated frames (KHR_FRAME_ACCUMULATION semantics): matches O-cpu's accumulated image (MAE)."""
    scene = H.default_scene(32, 64, 64, rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT,
                            fmt=capi.DVR_FORMAT_FLOAT32_VEC4)
    got = H.render_cuda(scene, frames=64)
    want = H.render_oracle(scene, frames=64)
    mae = float(np.abs(got["color"] - want["color"]).mean())
    assert mae < 0.5 / 255, mae
    one = H.render_cuda(scene, frames=1)
    # accumulation reduces the jitter noise: the 64-frame image is smoother than the 1-frame image
    lap = lambda img: np.abs(np.diff(img.reshape(64, 64, 4)[..., :3], axis=1)).mean()
    assert lap(got["color"]) < lap(one["color"])


@pytest.mark.skipif(not ob.have_ref_gpu(), reason="oracle/_ref/libref_gpu_dvr.so not built")
@pytest.mark.parametrize("bricks", ["1", "0"])
def test_config_c5_shape_64_frame_progressive_accumulation_on_a_nanovdb_fog(bricks, monkeypatch):
    """BASELINE config C5 as a parity case: a NanoVDB fog sphere (r = 30 voxels, our own writer), `default` renderer,
    64 frames of progressive accumulation (KHR_FRAME_ACCUMULATION) — the accumulated float frame against O-gpu (the
    reference's device code with the reference's NanoVDB sampler) after 64 frames: every pixel <= 2/255, MAE far below;
    apron bricks on and off (the tree walk) give the same frame bit for bit."""
    from visrtx_b200 import nvdb_writer
    monkeypatch.setenv("DVR_B200_NVDB_BRICKS", bricks)
    W, Hh = 160, 120
    blob = nvdb_writer.fog_sphere(30.0, voxel_size=1.0, half_width=3.0)
    v = H.VolumeDesc(np.zeros((1, 1, 1), np.float32), nvdb=blob, tf=capi.tf_discretize(
        color=scenes.tsd_default_colormap(256)), unit_distance=8.0, vol_id=5, inst_id=1)
    lo, hi = v.bounds()
    pose = scenes.orbit_camera(lo, hi, W, Hh, dist_scale=1.0)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    scene = H.SceneDesc([v], W, Hh, cam, volume_sampling_rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT,
                        fmt=capi.DVR_FORMAT_FLOAT32_VEC4)
    got = H.render_cuda(scene, frames=64, skip=True)
    want = H.render_refgpu(scene, frames=64)
    d = np.abs(got["color"] - want["color"]).max(axis=-1) * 255.0
    REPORT[f"live-O-gpu:c5-shape-64-frames-bricks{bricks}"] = {"max_abs_255": float(d.max()), "mae_255": float(d.mean())}
    assert d.max() <= 2.0 and d.mean() < 0.05, (float(d.max()), float(d.mean()))
    assert float(got["color"][..., 3].max()) > 0.5  # the fog is on screen
    one = H.render_cuda(scene, frames=1, skip=True)
    lap = lambda img: np.abs(np.diff(img.reshape(Hh, W, 4)[..., :3], axis=1)).mean()
    assert lap(got["color"]) < lap(one["color"])  # accumulation averaged the jitter noise
    if bricks == "0":
        monkeypatch.setenv("DVR_B200_NVDB_BRICKS", "1")
        again = H.render_cuda(scene, frames=64, skip=True)
        assert np.array_equal(again["color"], got["color"])


def test_float64_and_float16_fields():
    import torch
    base = scenes.blobs_np(24)
    s32 = H.default_scene(24, 64, 64, rate=0.5, field="blobs")
    s32.volumes[0].unit_distance = 0.5
    ref = H.render_cuda(s32)
    s64 = H.default_scene(24, 64, 64, rate=0.5, field="blobs")
    s64.volumes[0].unit_distance = 0.5
    s64.volumes[0].voxels = base.astype(np.float64)
    s64.volumes[0].data_type = capi.DVR_FLOAT64
    got = H.render_cuda(s64)
    assert np.array_equal(ref["color"], got["color"])
    s16 = H.default_scene(24, 64, 64, rate=0.5, field="blobs")
    s16.volumes[0].unit_distance = 0.5
    s16.volumes[0].voxels = base.astype(np.float16)
    s16.volumes[0].data_type = capi.DVR_FLOAT16
    got16 = H.render_cuda(s16)
    s16o = H.default_scene(24, 64, 64, rate=0.5, field="blobs")
    s16o.volumes[0].unit_distance = 0.5
    s16o.volumes[0].voxels = base.astype(np.float16).astype(np.float32)
    want16 = H.render_oracle(s16o)
    _check(got16, want16, s16, strict=False, name="f16", against="O-cpu")


def test_device_pointer_field_upload_equals_host_upload():
    """ANARI_NV_ARRAY_CUDA: a field created from a device pointer renders like one created from host memory."""
    import torch
    scene, _, _ = ZOO["ml48_raycast_r0.5"]
    a = H.render_cuda(scene)
    v = scene.volumes[0]
    dev = torch.from_numpy(v.voxels).cuda()
    f = capi.Field.create_structured(dev.data_ptr(), True, capi.DVR_FLOAT32, v.dims, v.origin, v.spacing)
    vol = capi.Volume.create(f, v.tf, v.value_range, v.unit_distance, v.vol_id)
    inst, n = capi.make_instances([vol], None, [v.inst_id])
    cs = H.CudaScene(scene)
    try:
        capi.render(H._params(scene, 0, -1), scene.camera, inst, n, cs.fb)
        b = cs.download()
    finally:
        cs.destroy()
        vol.destroy()
        f.destroy()
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_full_size_properties_c2():
    """BASELINE config 2 sizes (1024^3 f32, 1920x1080): properties that need no oracle run.
    idempotence (same frameID twice => identical bits), skipping == no skipping, opacity in [0,1],
    background where the volume is missed, sample count equals the sum over rays of lattice points."""
    import torch
    n, W, Hh = 1024, 1920, 1080
    vol = scenes.marschner_lobb_torch(n, "cuda")
    f = capi.Field.create_structured(vol.data_ptr(), True, capi.DVR_FLOAT32, (n, n, n), (0, 0, 0), (1, 1, 1))
    del vol
    tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256))
    v = capi.Volume.create(f, tf, (0.0, 1.0), 256.0, 0)
    inst, ni = capi.make_instances([v], None, [0])
    pose = scenes.orbit_camera((0, 0, 0), (n - 1,) * 3, W, Hh)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    npx = W * Hh
    accum = torch.empty((npx, 4), dtype=torch.float32, device="cuda")
    color = torch.empty((npx, 4), dtype=torch.float32, device="cuda")
    depth = torch.empty(npx, dtype=torch.float32, device="cuda")
    fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
    try:
        outs = []
        for skip in (False, False, True):
            p = capi.frame_params(W, Hh, capi.DVR_FORMAT_FLOAT32_VEC4, capi.DVR_INTEGRATOR_RAYCAST, 0, -1, 1, 0.5,
                                  (0.1, 0.1, 0.1, 1.0), skip=skip)
            capi.render(p, cam, inst, ni, fb)
            torch.cuda.synchronize()
            outs.append((color.clone(), depth.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert torch.equal(outs[0][0], outs[2][0])
        c, d = outs[0]
        miss = d >= 1e29
        assert 0.85 < miss.float().mean().item() < 0.97
        assert torch.allclose(c[miss][:, :3], torch.tensor([0.1, 0.1, 0.1], device="cuda"), atol=1e-6)
        assert c[:, 3].min().item() >= 0.999 and c[:, 3].max().item() <= 1.0 + 1e-5  # opaque background
        assert torch.isfinite(c).all()
        st = torch.zeros(4, dtype=torch.int64, device="cuda")
        capi.render_instrumented(p, cam, inst, ni, fb, st.data_ptr())
        torch.cuda.synchronize()
        taken, skipped, rays, cells = st.tolist()
        assert rays == int((~miss).sum().item())
        assert cells > 0.98 * 64 ** 3 and skipped == 0
        assert 300 < taken / rays < 1800  # chord lengths of a 1024^3 cube at 1 voxel per step
    finally:
        v.destroy()
        f.destroy()


@pytest.mark.parametrize("config", ["c2", "c3", "c5", "c5_tree_walk"])
def test_benchmark_configs_match_reference_device_code_at_full_size(config, monkeypatch):
    """The BASELINE.json configs AT THE SIZES THE BENCH TIMES — C2 1024^3 f32 @1080p, C3 2048^3 UFIXED16 sparse @4K
    with macrocell skipping, C5 NanoVDB fog r=200 @1080p (apron bricks, and the tree walk with bricks disabled) — frame
    0 of the benchmark camera against O-gpu (the reference's device code) live on this GPU: north_star's per-pixel
    <= 2/255 and PSNR >= 45 dB, depth to 2e-5 relative, identical hit masks.  Scene construction is bench.py's own."""
    if not ob.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu_dvr.so not present")
    import argparse
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    a = argparse.Namespace(config=config.split("_")[0], size=1024, width=1920, height=1080, rate=0.5, unit_distance=256.0,
                           field="ml", skip=-1, mode="auto", nvdb_codec="float", steps=1, warmup=0, cpu_rows=0)
    a = bench.apply_preset(a)
    if config == "c5_tree_walk":
        monkeypatch.setenv("DVR_B200_NVDB_BRICKS", "0")
    device = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    vol = bench.make_scene(a, torch, device)
    field = bench.create_field(a, capi, vol, stream)
    tf = capi.tf_discretize(color=bench.scene_colormap(a))
    v = capi.Volume.create(field, tf, (0.0, 1.0), a.unit_distance, 0, stream)
    inst, ninst = capi.make_instances([v], None, [0])
    cam, _ = bench.orbit(a)
    npx = a.width * a.height
    accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
    color = torch.zeros(npx, dtype=torch.int32, device=device)
    depth = torch.zeros(npx, dtype=torch.float32, device=device)
    fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
    try:
        p = capi.frame_params(a.width, a.height, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, 0, -1, 1,
                              a.rate, (0.1, 0.1, 0.1, 1.0), skip=bool(a.skip))
        capi.render(p, cam, inst, ninst, fb, stream)
        torch.cuda.synchronize()
        (ref_color, ref_depth), _ = bench.ref_gpu_frame0(a, torch, vol)
        par = bench.parity_block(torch, 4, color, ref_color, depth, ref_depth, what=config)
        REPORT[f"live-O-gpu:full-size-{config}"] = par
        assert par["max_abs_255"] <= 2 and par["psnr_db"] >= 45.0, par
        assert par["depth_hit_mask_equal"] and par["depth_max_rel"] <= 2e-5, par
        assert (ref_depth < 1e29).float().mean().item() > 0.02  # the volume is on screen
    finally:
        v.destroy()
        field.destroy()


def test_closed_form_lattice_advance_is_bit_identical_to_the_add_loop():
    """4 M pseudo-random (t, step, n, tUpper) sets, incl. power-of-two steps (round-to-even ties), tiny and huge t."""
    for seed in (1, 77, 2024, 999983):
        assert capi.selftest_lattice_advance(1 << 20, seed) == 0


def test_automatic_skipping_mode_picks_by_emptiness_and_never_changes_pixels():
    dense = H.default_scene(48, 96, 64, rate=0.5)  # Marschner-Lobb + default map: no empty macrocell
    sparse = H.default_scene(64, 96, 64, rate=0.5)
    vox = np.zeros((64, 64, 64), np.float32)
    vox[20:44, 20:44, 20:44] = scenes.blobs_np(24)  # a blob in the middle of empty space
    sparse.volumes[0].voxels = vox
    sparse.volumes[0].tf = capi.tf_discretize(color=scenes.sparse_colormap(256, 0.2))
    for scene, expect_skips in ((dense, False), (sparse, True)):
        cs = H.CudaScene(scene)
        try:
            p = H._params(scene, 0, -1, skip="auto")
            assert p.useMacrocellSkipping == capi.DVR_SKIP_AUTO
            import torch
            st = torch.zeros(4, dtype=torch.int64, device="cuda")
            capi.render_instrumented(p, scene.camera, cs.instances, cs.n, cs.fb, st.data_ptr())
            torch.cuda.synchronize()
            auto = cs.download()
            skipped = int(st[1].item())
        finally:
            cs.destroy()
        assert (skipped > 0) == expect_skips
        ref = H.render_cuda(scene, skip=False)
        for k in ref:
            assert np.array_equal(auto[k], ref[k]), k


def test_partial_march_confined_to_the_screen_rectangle_of_the_bounds():
    """partialCullToBounds: tiles outside the projected box are neither marched nor written.  Every pixel whose ray
    hits the volume must still be produced (conservative rectangle): compared with the full-frame partial on
    cameras outside / far / cropped / orthographic / inside (inside and thin-lens fall back to the whole frame)."""
    import torch
    n = 32
    vox = scenes.blobs_np(n)
    W, Hh = 200, 144
    cams = []
    for az, el, dist in ((30, 20, 2.0), (200, -35, 1.2), (95, 5, 3.5), (10, 80, 1.6)):
        v = H.default_scene(n, W, Hh, field="blobs").volumes[0]
        lo, hi = v.bounds()
        pose = scenes.orbit_camera(lo, hi, W, Hh, az_deg=az, el_deg=el, dist_scale=dist)
        cams.append(capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect))
        cams.append(capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect,
                                            region=(0.2, 0.1, 0.9, 0.75)))
    cams.append(capi.camera_orthographic((0.3, 0.2, 5.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 4.0, W / Hh))
    cams.append(capi.camera_orthographic((4.0, 3.0, 5.0), (-0.5, -0.4, -0.7), (0.0, 1.0, 0.0), 1.5, W / Hh))
    cams.append(capi.camera_perspective((0.1, 0.0, 0.2), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 1.0, W / Hh))  # inside
    cams.append(capi.camera_perspective((0.0, 0.0, 4.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 0.6, W / Hh, 4.0, 0.1))  # lens
    cams.append(capi.camera_perspective((0.0, 0.0, 4.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.6, W / Hh))  # looking away
    npx = W * Hh
    culled_any = False
    for ci, cam in enumerate(cams):
        scene = H.default_scene(n, W, Hh, rate=0.5, field="blobs", integrator=capi.DVR_INTEGRATOR_DEFAULT)
        scene.camera = cam
        cs = H.CudaScene(scene)
        try:
            full_rgba = torch.zeros((npx, 4), dtype=torch.float32, device="cuda")
            full_depth = torch.zeros(npx, dtype=torch.float32, device="cuda")
            capi.render_partial(H._params(scene, 0, -1), cam, cs.instances, full_rgba.data_ptr(), full_depth.data_ptr())
            cut_rgba = torch.full((npx, 4), float("nan"), dtype=torch.float32, device="cuda")
            cut_depth = torch.full((npx,), float("nan"), dtype=torch.float32, device="cuda")
            capi.render_partial(H._params(scene, 0, -1, partial_cull_to_bounds=True), cam, cs.instances,
                                cut_rgba.data_ptr(), cut_depth.data_ptr())
            torch.cuda.synchronize()
        finally:
            cs.destroy()
        fr, fd = full_rgba.cpu().numpy(), full_depth.cpu().numpy()
        cr, cd = cut_rgba.cpu().numpy(), cut_depth.cpu().numpy()
        written = ~np.isnan(cd)
        hit = fd < 1e29  # pixels whose ray entered the bounds
        assert np.all(written[hit]), (ci, int((hit & ~written).sum()))
        assert np.array_equal(cr[written], fr[written]) and np.array_equal(cd[written], fd[written]), ci
        culled_any |= bool((~written).any())
    assert culled_any  # the far cameras really skip most of the frame


def _field_ranges(f):
    import ctypes
    import torch
    (gx, gy, gz), ptr = f.macrocells()
    rng = torch.empty((gz, gy, gx, 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(rng.data_ptr()), ctypes.c_void_p(ptr),
                                           ctypes.c_size_t(rng.numel() * 4), ctypes.c_int(3))
    return rng.cpu().numpy()


@pytest.mark.parametrize("dims", [(70, 33, 50), (128, 40, 37), (260, 19, 65), (16, 16, 16), (4, 1, 1), (516, 35, 34)])
def test_macrocell_build_from_device_memory_matches_texture_build(dims):
    """K4 has two builds: through the point-sampled array (host data, fixed point, slabs) and the separable one straight
    from linear f32 device memory (in-situ / time-varying fields).  Same ranges, bit for bit, NaNs dropped alike;
    dims cover the scalar x pass (x % 4 != 0), the vectorised one, several 128-voxel segments and ragged last cells."""
    import torch
    nx, ny, nz = dims
    g = np.random.default_rng(nx * 7 + ny)
    vox = g.standard_normal((nz, ny, nx)).astype(np.float32)
    if vox.size > 64:
        vox.reshape(-1)[g.integers(0, vox.size, 5)] = np.nan
    dev = torch.from_numpy(vox).cuda()
    a = capi.Field.create_structured(vox.ctypes.data, False, capi.DVR_FLOAT32, dims, (0, 0, 0), (1, 1, 1))
    b = capi.Field.create_structured(dev.data_ptr(), True, capi.DVR_FLOAT32, dims, (0, 0, 0), (1, 1, 1))
    # an unaligned device pointer takes the scalar x pass
    pad = torch.empty(vox.size + 1, dtype=torch.float32, device="cuda")
    pad[1:] = dev.reshape(-1)
    c = capi.Field.create_structured(pad.data_ptr() + 4, True, capi.DVR_FLOAT32, dims, (0, 0, 0), (1, 1, 1))
    try:
        ra, rb, rc = _field_ranges(a), _field_ranges(b), _field_ranges(c)
        assert ra.shape == ((nz + 15) // 16, (ny + 15) // 16, (nx + 15) // 16, 2)
        assert np.array_equal(ra, rb) and np.array_equal(ra, rc)
        for cz, cy, cx in np.ndindex(*ra.shape[:3]):
            blk = vox[max(cz * 16 - 1, 0):cz * 16 + 18, max(cy * 16 - 1, 0):cy * 16 + 18,
                      max(cx * 16 - 1, 0):cx * 16 + 18]
            assert rb[cz, cy, cx, 0] == np.nanmin(blk) and rb[cz, cy, cx, 1] == np.nanmax(blk)
        assert a.value_range() == b.value_range()
    finally:
        a.destroy()
        b.destroy()
        c.destroy()


@pytest.mark.parametrize("name", ["nvdb_fog_r20", "nvdb_fog_r12_vs025", "nvdb_fp4_r14", "nvdb_fp8_r14", "nvdb_fp16_r14",
                                  "nvdb_fpn_r14"])
@pytest.mark.parametrize("skip", [False, True])
def test_nanovdb_apron_bricks_are_bit_identical_to_the_tree_walk(name, skip, monkeypatch):
    """NanoVDB fields are gathered into 9^3 apron bricks at creation (dvr_nvdb_bricks.cu) and sampled from those;
    DVR_B200_NVDB_BRICKS=0 keeps the per-tap tree walk that mirrors the reference's accessor.  Same voxel values,
    same interpolation arithmetic: every channel of the frame is identical, for float and quantised grids, with
    negative index coordinates, a non-unit voxel size and an off-origin grid."""
    scene, frames, cb = ZOO[name]
    a = H.render_cuda(scene, frames=frames, checkerboard=cb, skip=skip)
    monkeypatch.setenv("DVR_B200_NVDB_BRICKS", "0")
    b = H.render_cuda(scene, frames=frames, checkerboard=cb, skip=skip)
    for k in a:
        assert np.array_equal(a[k], b[k]), (name, k)


def test_nanovdb_apron_bricks_cover_the_grid_and_count_as_device_memory(monkeypatch):
    from visrtx_b200 import nvdb_writer
    grid = nvdb_writer.fog_sphere(radius=20.0, voxel_size=1.0, half_width=3.0)
    f = capi.Field.create_nanovdb(grid.ctypes.data, grid.nbytes)
    monkeypatch.setenv("DVR_B200_NVDB_BRICKS", "0")
    g = capi.Field.create_nanovdb(grid.ctypes.data, grid.nbytes)
    try:
        extra = f.device_bytes() - g.device_bytes()
        # table (8 B per 8^3 cell of the bounding box) + at least the bricks of the sphere's shell, at most one per cell
        cells = 6 ** 3  # index bounding box [-19, 19]^3: cells (-20 >> 3) .. (19 >> 3) per axis
        assert extra > cells * 8 and extra <= cells * (8 + 729 * 4)
    finally:
        f.destroy()
        g.destroy()
