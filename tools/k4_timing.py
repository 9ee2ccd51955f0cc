"""Times K4 (macrocell value ranges) and the in-place field refresh at BASELINE config-2 size on the GPU.

  texture build   dvr_field_build_macrocells: one CTA per cell through the point-sampled 3-D array
  refresh         dvr_field_update_structured: ONE pass over linear f32 device memory that stores the voxels into the
                  3-D array through its surface and reduces the macrocell ranges (DVR_B200_SURFACE_UPLOAD=0: the
                  driver's array copy followed by the separable build), then dvr_volume_update
  re-create       what the reference does on a field commit: destroy, cudaMalloc3DArray, copy, build

usage: python tools/k4_timing.py [--size 1024] [--reps 5]   -> one JSON line
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from visrtx_b200 import capi, scenes
    n = a.size
    vol = scenes.marschner_lobb_torch(n, "cuda")
    tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256))
    f = capi.Field.create_structured(vol.data_ptr(), True, capi.DVR_FLOAT32, (n, n, n), (0, 0, 0), (1, 1, 1))
    v = capi.Volume.create(f, tf, (0.0, 1.0), 256.0, 0)
    torch.cuda.synchronize()

    def timed(fn):
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    def wall(fn):
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t) * 1e3)
        return sorted(ts)[len(ts) // 2]

    out = {"size": n, "voxel_GB": n ** 3 * 4 / 1e9}
    out["texture_build_ms"] = timed(lambda: f.build_macrocells(0))
    out["refresh_field_ms"] = timed(lambda: f.update_structured(vol.data_ptr(), True, capi.DVR_FLOAT32, (0, 0, 0), (1, 1, 1)))
    # copy alone, to separate the linear build from the array upload
    out["array_copy_ms"] = timed(lambda: f.upload_slices(vol.data_ptr(), True, 0, n, 0))
    out["refresh_field_GBps_of_input"] = out["voxel_GB"] / (out["refresh_field_ms"] * 1e-3)
    out["surface_upload"] = os.environ.get("DVR_B200_SURFACE_UPLOAD", "1") != "0"

    def refresh():
        f.update_structured(vol.data_ptr(), True, capi.DVR_FLOAT32, (0, 0, 0), (1, 1, 1))
        v.update(tf, (0.0, 1.0), 256.0, 0)

    out["refresh_field_and_volume_wall_ms"] = wall(refresh)
    state = {"f": None, "v": None}

    def recreate():
        if state["v"]:
            state["v"].destroy()
            state["f"].destroy()
        state["f"] = capi.Field.create_structured(vol.data_ptr(), True, capi.DVR_FLOAT32, (n, n, n), (0, 0, 0), (1, 1, 1))
        state["v"] = capi.Volume.create(state["f"], tf, (0.0, 1.0), 256.0, 0)

    v.destroy()
    f.destroy()
    out["recreate_field_and_volume_wall_ms"] = wall(recreate)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
