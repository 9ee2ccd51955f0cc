"""The plain-C ANARI application in examples/render_volume_file.c: the public headers compile as C99, the program
links against the drop-in libraries only, fails loudly without a GPU (CPU suite) and renders the imported file to
the same pixels as the Python-driven path (GPU suite)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "visrtx_b200")


def _build(tmp_path):
    exe = str(tmp_path / "render_volume_file")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "render_volume_file.c"), "-L" + LIBDIR, "-lanari_library_visrtx_b200",
           "-ldvr_import", "-ldvr_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath-link," + LIBDIR, "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def _raw(tmp_path, n=24):
    from visrtx_b200 import scenes
    vox = (scenes.blobs_np(n) * 255).astype(np.uint8)
    p = tmp_path / f"blobs_{n}x{n}x{n}_uint8.raw"
    vox.tofile(p)
    return str(p), vox


def test_example_compiles_as_c99_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("this half of the test is for GPU-less machines")
    raw, _ = _raw(tmp_path)
    r = subprocess.run([exe, raw, str(tmp_path / "o.ppm"), "raycast", "1", "32", "32"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "o.ppm")


@pytest.mark.gpu
def test_example_renders_the_same_pixels_as_the_python_path(tmp_path):
    import dvr_harness as H
    from visrtx_b200 import capi, importers, scenes
    exe = _build(tmp_path)
    raw, vox = _raw(tmp_path)
    out = tmp_path / "o.ppm"
    r = subprocess.run([exe, raw, str(out), "default", "3", "96", "64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "structuredRegular 24x24x24" in r.stdout
    data = out.read_bytes()
    assert data.startswith(b"P6\n96 64\n255\n")
    img = np.frombuffer(data[len(b"P6\n96 64\n255\n"):], np.uint8).reshape(64, 96, 3)[::-1]  # undo the PPM flip
    # the same scene through the C-ABI harness: import_volume semantics (default colour map, valueRange = data range)
    vf = importers.import_raw(raw)
    cmap = np.array([[1, 0, 0, 0], [0, 1, 0, .5], [0, 0, 1, 1]], np.float32)
    v = H.VolumeDesc(vox, data_type=capi.DVR_UFIXED8, origin=(0.0,) * 3, spacing=(1.0,) * 3, value_range=vf.value_range,
                     tf=capi.tf_discretize(color=cmap, value_range=vf.value_range), unit_distance=4.0)
    lo, hi = v.bounds()
    pose = scenes.orbit_camera(lo, hi, 96, 64, dist_scale=1.2)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    scene = H.SceneDesc([v], 96, 64, cam, volume_sampling_rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT)
    ref = H.unpack_rgba8(H.render_cuda(scene, frames=3)["color"]).reshape(64, 96, 4)[..., :3]
    d = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    # the C program computes its camera in float with sinf/cosf, the Python path in double: rays differ by ulps
    assert (d <= 2).mean() >= 0.995 and d.max() <= 12, (float((d <= 2).mean()), int(d.max()))
