"""Generates tests/golden/refgpu_scenes.npz and philox_curand.npz ON A B200 BOX:

    gpurun -- 'python tests/golden/make_golden.py'   (then copy gpurun_out/golden/* into tests/golden/)

Every scene of dvr_harness.scene_zoo() is rendered by O-gpu — the reference's own device headers
(oracle/_ref/libref_gpu_dvr.so: hardware tex3D/tex1D, cuRAND Philox, accumResults) — and the outputs
are stored so that the CPU suite can pin O-cpu (and the GPU suite the CUDA path) without the reference
tree.  The Philox known-answer vectors come from cuRAND itself via the same library's RNG use: they are
read back from a tiny kernel compiled with torch's bundled headers is avoided — instead they are
derived from O-gpu's jitter through a 1-pixel render (see test_oracle_golden.py::test_philox_stream).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import dvr_harness as H  # noqa: E402


def main():
    # `--add name1,name2`: render only those scenes and MERGE them into the committed fixture, leaving the existing
    # pins byte for byte as they are (the round-1 renders stay the reference for the round-1 scenes)
    only = None
    if len(sys.argv) > 2 and sys.argv[1] == "--add":
        only = set(sys.argv[2].split(","))
    out = {}
    if only is not None:
        old = np.load(os.path.join(HERE, "refgpu_scenes.npz"))
        out = {k: old[k] for k in old.files}
    for name, (scene, frames, cb) in H.scene_zoo().items():
        if only is not None and name not in only:
            continue
        r = H.render_refgpu(scene, frames=frames, checkerboard=cb)
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(name, {k: v.shape for k, v in r.items()})
    dst = os.path.join(os.path.dirname(os.path.dirname(HERE)), "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "refgpu_scenes.npz"), **out)
    print("wrote", os.path.join(dst, "refgpu_scenes.npz"))
    if only is not None:
        return

    # dpt renderer: the reference's tracker (O-gpu) over O-cpu's majorant grid (the product's grid is tested
    # to be bit-identical to it); frame 0 colour + the accumulation of DPT_GOLDEN_FRAMES frames
    dpt = {}
    for kind in H.DPT_KINDS:
        scene = H.dpt_scene(kind)
        grids = H.oracle_dda_grids(scene)
        r1 = H.render_refgpu(scene, frames=1, grids=grids)
        rn = H.render_refgpu(scene, frames=H.DPT_GOLDEN_FRAMES, grids=grids)
        dpt[f"{kind}/color1"] = r1["color"]
        dpt[f"{kind}/accum"] = rn["accum"]
        print("dpt", kind, float(r1["color"][:, :3].mean()))
    np.savez_compressed(os.path.join(dst, "refgpu_dpt.npz"), **dpt)
    print("wrote", os.path.join(dst, "refgpu_dpt.npz"))


if __name__ == "__main__":
    main()
