"""GPU suite, multi-GPU: sort-first (bit-identical) and sort-last (<= 1/255) against the single-GPU frame,
run as real one-process-per-GPU jobs over NCCL + CUDA IPC when the box has >= 2 GPUs; the world-size-1
degenerate path of both drivers always runs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import dvr_harness as H
from visrtx_b200 import capi, multigpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_drivers_world_size_one():
    import torch
    scene = H.default_scene(48, 120, 88, rate=0.5, field="blobs", integrator=capi.DVR_INTEGRATOR_DEFAULT)
    scene.volumes[0].unit_distance = 0.5
    v = scene.volumes[0]
    single = H.render_cuda(scene, frames=2)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    cs = H.CudaScene(scene)
    sf = multigpu.SortFirst(capi, torch, None, 0, 1, dev, scene.width, scene.height, cs.instances, cs.n, scene.fmt,
                            scene.integrator, scene.volume_sampling_rate, scene.background)
    sl = multigpu.SortLast(capi, torch, None, 0, 1, dev, scene.width, scene.height, cs.instances, v.vol_id, v.inst_id,
                           scene.fmt, scene.integrator, scene.volume_sampling_rate, scene.background)
    for fid in range(2):
        sf.render(fid, scene.camera, stream)
        sl.render(fid, scene.camera, stream)
    torch.cuda.synchronize()
    a = sf.color_tensor().cpu().numpy().view(np.uint32)
    b = sl.color_tensor().cpu().numpy().view(np.uint32)
    assert np.array_equal(a, single["color"])
    d = np.abs(H.unpack_rgba8(b) - H.unpack_rgba8(single["color"])).max(axis=-1)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 3
    sf.close()
    sl.close()
    cs.destroy()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_process_sort_first_and_sort_last(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MGPU_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
