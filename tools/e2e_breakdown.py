"""Where the end-to-end step time goes (C2 through the ANARI C API): wall-clock per call, averaged."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    sys.argv = ["bench.py"] + sys.argv[1:]
    args = bench.parse_args()
    device = torch.device("cuda", 0)
    vol = bench.make_scene(args, torch, device)
    e = bench.AnariE2E(args, torch, device, vol, 0, 1, "single")
    n = 200
    e.prepare(n + 5)
    for i in range(5):
        e.step(i)
    A, d = e.A, e.d
    t = {"set+commit": 0.0, "render": 0.0, "ready": 0.0, "map": 0.0, "unmap": 0.0}
    dur = 0.0
    for i in range(n):
        t0 = time.perf_counter()
        for name, dt, _keep, ptr in e.inputs[i]:
            A.lib.anariSetParameter(d.handle, e.camera, name, dt, ptr)
        A.lib.anariCommitParameters(d.handle, e.camera)
        t1 = time.perf_counter()
        A.lib.anariRenderFrame(d.handle, e.frame)
        t2 = time.perf_counter()
        A.lib.anariFrameReady(d.handle, e.frame, A.WAIT)
        t3 = time.perf_counter()
        p = A.lib.anariMapFrame(d.handle, e.frame, b"channel.color", C.byref(e.w), C.byref(e.h), C.byref(e.t))
        e.checksum ^= C.cast(p, C.POINTER(C.c_uint32))[100]
        t4 = time.perf_counter()
        A.lib.anariUnmapFrame(d.handle, e.frame, b"channel.color")
        t5 = time.perf_counter()
        for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            t[k] += v
        dur += d.get_property(e.frame, "duration", A.FLOAT32)
    print({k: round(v / n * 1e6, 1) for k, v in t.items()}, "us per step;", "device duration",
          round(dur / n * 1e6, 1), "us; total", round(sum(t.values()) / n * 1e6, 1), "us")
    e.close()


if __name__ == "__main__":
    main()
