"""Renders a few frames of one BASELINE scene through the C-ABI and nothing else — the short command ncu wraps
(B200_PROFILING.md: keep the profiled command short).  Scenes and cameras are bench.py's own.

    python tools/profile_scene.py --what c5 --frames 12          # NanoVDB fog, apron-brick march
    python tools/profile_scene.py --what dpt --frames 12         # delta tracking on the C2 field
    python tools/profile_scene.py --what c2 --rate 1.0           # C2 at half a voxel per step
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="c2", choices=["c2", "c3", "c5", "dpt"])
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--rate", type=float, default=0.5)
    ap.add_argument("--nvdb-codec", default="float")
    o = ap.parse_args()
    cfg = "c2" if o.what == "dpt" else o.what
    sys.argv = ["bench.py", "--config", cfg, "--rate", str(o.rate), "--nvdb-codec", o.nvdb_codec]
    import bench
    import numpy as np
    import torch
    from visrtx_b200 import capi
    a = bench.parse_args()
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    stream = torch.cuda.current_stream().cuda_stream
    vol = bench.make_scene(a, torch, device)
    field = bench.create_field(a, capi, vol, stream)
    torch.cuda.synchronize()
    integ = capi.DVR_INTEGRATOR_DEFAULT
    if o.what == "dpt":
        tf = capi.tf_discretize(color=bench.scene_colormap(a), opacity=np.asarray(bench.DPT_OPACITY, np.float32))
        integ = capi.DVR_INTEGRATOR_DPT
    else:
        tf = capi.tf_discretize(color=bench.scene_colormap(a))
    v = capi.Volume.create(field, tf, (0.0, 1.0), a.unit_distance, 0, stream)
    inst, ninst = capi.make_instances([v], None, [0])
    cam, _ = bench.orbit(a)
    npx = a.width * a.height
    accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
    color = torch.zeros(npx, dtype=torch.int32, device=device)
    depth = torch.zeros(npx, dtype=torch.float32, device=device)
    fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(o.frames):
        if i == o.frames // 2:
            e0.record()
        p = capi.frame_params(a.width, a.height, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, integ, i, -1, 1, a.rate,
                              (0.1, 0.1, 0.1, 1.0), skip=bool(a.skip), max_depth=5, ambient_radiance=1.0)
        capi.render(p, cam, inst, ninst, fb, stream)
    e1.record()
    torch.cuda.synchronize()
    n = o.frames - o.frames // 2
    print(f"{o.what}: {bench.workload_name(a)}: {e0.elapsed_time(e1) / n:.4f} ms per frame over the last {n} frames")


if __name__ == "__main__":
    main()
