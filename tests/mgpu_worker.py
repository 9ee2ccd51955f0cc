"""Multi-rank worker (launched by torchrun): sort-first and sort-last renders of a test scene on N GPUs,
compared on rank 0 with the single-GPU frame.  Used by tests/test_gpu_multigpu.py and by hand:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dvr_harness as H  # noqa: E402
from visrtx_b200 import capi, multigpu  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    capi.set_device(local)
    dist.init_process_group("nccl")
    device = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    ok = True

    scene = H.default_scene(64, 200, 152, rate=0.5, field="blobs", integrator=capi.DVR_INTEGRATOR_DEFAULT)
    scene.volumes[0].unit_distance = 0.5
    v = scene.volumes[0]
    single = H.render_cuda(scene, frames=3) if rank == 0 else None

    # ---- sort-first: replicated field, interleaved tile rows, peer stores into rank 0's frame
    cs = H.CudaScene(scene, device=f"cuda:{local}")
    sf = multigpu.SortFirst(capi, torch, dist, rank, world, device, scene.width, scene.height, cs.instances, cs.n,
                            scene.fmt, scene.integrator, scene.volume_sampling_rate, scene.background, host_mirror=True)
    for fid in range(3):
        sf.render(fid, scene.camera, stream)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        got = sf.color_tensor().cpu().numpy().view(np.uint32)
        same = np.array_equal(got, single["color"])
        print(f"[sort-first x{world}] bit-identical to single GPU: {same}")
        ok &= same
        # the shared pinned host frame every rank streamed its tile rows into holds the same image
        same_host = np.array_equal(np.array(sf.host_frame.numpy(), copy=True), got)
        print(f"[sort-first x{world}] shared host frame == device frame: {same_host}")
        ok &= same_host
    sf.close()
    cs.destroy()

    # ---- sort-last: z-slabs on the global lattice.  fused: one launch per GPU and frame (march + region flags +
    # composite/resolve over peer memory); legacy: partial march, wait, peer composite, signal
    nz = v.dims[2]
    z0, z1 = multigpu.slab_ranges(nz, world)[rank]
    for fused in (True, False):
        cs = H.CudaScene(scene, slab=(z0, z1) if world > 1 else None, device=f"cuda:{local}")
        sl = multigpu.SortLast(capi, torch, dist, rank, world, device, scene.width, scene.height, cs.instances, v.vol_id,
                               v.inst_id, scene.fmt, scene.integrator, scene.volume_sampling_rate, scene.background,
                               host_mirror=True, fused=fused)
        for fid in range(3):
            sl.render(fid, scene.camera, stream)
        torch.cuda.synchronize()
        dist.barrier()
        err = torch.tensor([1 if sl.check_errors() else 0], device=device)
        dist.all_reduce(err)
        if rank == 0:
            tag = f"[sort-last x{world} {'fused' if fused else 'legacy'}]"
            print(f"{tag} bounded spins that gave up: {int(err.item())}")
            ok &= int(err.item()) == 0
            got = sl.color_tensor().cpu().numpy().view(np.uint32)
            same_host = np.array_equal(np.array(sl.host_frame.numpy(), copy=True), got)
            print(f"{tag} shared host frame == device frame: {same_host}")
            ok &= same_host
            d = np.abs(H.unpack_rgba8(got) - H.unpack_rgba8(single["color"])).max(axis=-1)
            good = (d <= 1).mean() >= 0.999 and d.max() <= 2
            print(f"{tag} max diff {d.max()}/255, frac<=1/255 {(d <= 1).mean():.5f}: {good}")
            ok &= bool(good)
            depth = sl.depth_tensor().cpu().numpy()
            dgood = np.allclose(depth, single["depth"], rtol=1e-6, atol=0)
            print(f"{tag} assembled depth == single-GPU depth: {dgood}")
            ok &= bool(dgood)
        sl.close()
        cs.destroy()

    # ---- the very first frame of a fresh driver: every pixel of the display frame (volume window AND background strips
    # of every rank) must be written by that one launch — nothing may rely on what an earlier frame left behind
    if world > 1:
        single0 = H.render_cuda(scene, frames=1) if rank == 0 else None
        for fused in (True, False):
            cs = H.CudaScene(scene, slab=(z0, z1), device=f"cuda:{local}")
            sl = multigpu.SortLast(capi, torch, dist, rank, world, device, scene.width, scene.height, cs.instances, v.vol_id,
                                   v.inst_id, scene.fmt, scene.integrator, scene.volume_sampling_rate, scene.background,
                                   fused=fused)
            if rank == 0:  # poison the display frame
                import ctypes
                px_bytes = 4
                ctypes.CDLL("libcudart.so").cudaMemset(ctypes.c_void_p(sl.color_ptr), 0xAB,
                                                       ctypes.c_size_t(sl.npx * px_bytes + sl.npx * 4))
            torch.cuda.synchronize()
            dist.barrier()
            sl.render(0, scene.camera, stream)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                got = sl.color_tensor().cpu().numpy().view(np.uint32)
                d = np.abs(H.unpack_rgba8(got) - H.unpack_rgba8(single0["color"])).max(axis=-1)
                good = (d <= 1).mean() >= 0.999 and d.max() <= 2
                print(f"[sort-last x{world} {'fused' if fused else 'legacy'}, first frame of a fresh driver] max diff "
                      f"{d.max()}/255, pixels off by > 2: {int((d > 2).sum())}: {good}")
                ok &= bool(good)
                depth = sl.depth_tensor().cpu().numpy()
                dgood = np.allclose(depth, single0["depth"], rtol=1e-6, atol=0)
                print(f"[sort-last x{world} {'fused' if fused else 'legacy'}, first frame] depth == single-GPU depth: {dgood}")
                ok &= bool(dgood)
            dist.barrier()
            sl.close()
            cs.destroy()

    # ---- ownership shift (sort-last load balancing): slabs created with a margin of resident slices, cuts moved with
    # dvr_field_set_owned_slices (no voxel moves) — by hand, then by the measured-time feedback of SortLast.calibrate
    if world > 1:
        single3 = H.render_cuda(scene, frames=3) if rank == 0 else None
        ranges0 = multigpu.slab_ranges(nz, world)
        margin = 5
        limits = multigpu.creation_ranges(ranges0, nz, margin)
        cs = H.CudaScene(scene, slab=limits[rank], device=f"cuda:{local}")
        fld = cs.fields[0]
        fld.set_owned_slices(*ranges0[rank])
        sl = multigpu.SortLast(capi, torch, dist, rank, world, device, scene.width, scene.height, cs.instances, v.vol_id,
                               v.inst_id, scene.fmt, scene.integrator, scene.volume_sampling_rate, scene.background)
        shifted = [(a + (2 if i > 0 else 0) * (1 if i % 2 else -1), b + (2 if i < world - 1 else 0) * (-1 if i % 2 else 1))
                   for i, (a, b) in enumerate(ranges0)]  # cut i moves by +-2 slices, alternating
        for label, rng in (("initial cuts", ranges0), ("cuts moved by 2 slices", shifted), ("calibrated", None)):
            if rng is None:
                rng, times, hist = sl.calibrate(fld, shifted, limits, scene.camera, stream, rounds=2, frames=3)
                if rank == 0:
                    print(f"[sort-last x{world} balance] calibrated cuts {rng} from march times {[round(t, 3) for t in times]} ms")
            else:
                fld.set_owned_slices(*rng[rank])
            assert fld.owned_slices() == (tuple(rng[rank]), tuple(limits[rank]))
            for fid in range(3):
                sl.render(fid, scene.camera, stream)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                got = sl.color_tensor().cpu().numpy().view(np.uint32)
                d = np.abs(H.unpack_rgba8(got) - H.unpack_rgba8(single3["color"])).max(axis=-1)
                good = (d <= 1).mean() >= 0.999 and d.max() <= 2
                print(f"[sort-last x{world} balance, {label}] max diff {d.max()}/255, frac<=1/255 {(d <= 1).mean():.5f}: {good}")
                ok &= bool(good)
            dist.barrier()
        bad = False
        try:
            fld.set_owned_slices(limits[rank][0], limits[rank][1] + 1)
        except capi.DvrError:
            bad = True
        ok &= bad
        sl.close()
        cs.destroy()

    # ---- a moving camera (accumulation reset every frame, the screen window changes) and a camera that does not see
    # the volume at all (empty window: background strips only)
    if world > 1:
        from visrtx_b200 import scenes
        cs = H.CudaScene(scene, slab=(z0, z1), device=f"cuda:{local}")
        sl = multigpu.SortLast(capi, torch, dist, rank, world, device, scene.width, scene.height, cs.instances, v.vol_id,
                               v.inst_id, scene.fmt, scene.integrator, scene.volume_sampling_rate, scene.background)
        lo, hi = v.bounds()
        cams = []
        for az in (10.0, 75.0, 160.0, 250.0):
            pose = scenes.orbit_camera(lo, hi, scene.width, scene.height, az_deg=az, el_deg=-15.0, dist_scale=0.8)
            cams.append(capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect))
        cams.append(capi.camera_perspective((0.0, 0.0, 9.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.8, scene.width / scene.height))
        for i, cam_i in enumerate(cams):
            sl.render(0, cam_i, stream)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                got = sl.color_tensor().cpu().numpy().view(np.uint32)
                moved = H.SceneDesc(scene.volumes, scene.width, scene.height, cam_i, fmt=scene.fmt, integrator=scene.integrator,
                                    volume_sampling_rate=scene.volume_sampling_rate, background=scene.background)
                want = H.render_cuda(moved)
                d = np.abs(H.unpack_rgba8(got) - H.unpack_rgba8(want["color"])).max(axis=-1)
                good = (d <= 1).mean() >= 0.999 and d.max() <= 2
                print(f"[sort-last x{world} fused, camera {i}] max diff {d.max()}/255: {good}")
                ok &= bool(good)
            dist.barrier()
        sl.close()
        cs.destroy()

    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_WORKER_OK" if ok else "MGPU_WORKER_FAILED")
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
