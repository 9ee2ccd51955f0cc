#!/usr/bin/env python
"""bench.py — DVR frames/s on BASELINE config C2 (1024^3 f32 volume, 1920x1080), B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one frame: one pass of the DVR hot path (ray generation, march, background,
accumulate/tonemap/encode) over all W*H pixels of a synthetic analytic volume that is already
resident in HBM.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte model.

  value     frames/s, device-timed (CUDA events on the launching stream), progressive accumulation
  e2e       frames/s through the public API with HOST buffers: every step re-commits a moved camera
            (host -> device parameter upload), renders, and maps the colour channel to pinned host
            memory (device -> host copy) inside the timed region
  roofline  HBM: algorithmic bytes per frame (voxels of every macrocell the rays sample, measured
            by an untimed instrumented launch, + 44 B/pixel + the 4 KiB TF) / average kernel time
  cpu_baseline  O-cpu (oracle/liboracle_dvr.so) on a bounded band of image rows, all host cores

--impl reference times O-gpu: the reference's OWN device headers (volumeIntegration.h etc.) compiled
for sm_100a with an OptiX shim (oracle/_ref/libref_gpu_dvr.so) — VisRTX itself cannot be built
without OptiX/ANARI-SDK — falling back to the O-cpu port when that library is absent.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# ---------------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="volume edge length (C2: 1024)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--rate", type=float, default=0.5, help="volumeSamplingRate (0.5 => 1 voxel per step)")
    ap.add_argument("--unit-distance", type=float, default=256.0,
                    help="TF unitDistance in voxels; 256 = semi-transparent, every ray crosses the volume")
    ap.add_argument("--field", default="ml", choices=["ml", "shells", "fog"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json config preset: c2 = 1024^3 f32 1080p (headline); c3 = 2048^3 16-bit sparse shells, "
                         "3840x2160, macrocell skipping; c4 = 4096^3 f32 sort-last (needs >= 2 GPUs)")
    ap.add_argument("--skip", type=int, default=-1,
                    help="macrocell skipping: 1/0; default -1 = off for the dense C2/C4 fields, on for C3")
    ap.add_argument("--mode", default="auto", choices=["auto", "sort-first", "sort-last"])
    ap.add_argument("--fused", type=int, default=1,
                    help="sort-last, N > 1: 1 = one fused launch per GPU and frame (march + exchange + composite), "
                         "0 = the round-1 sequence (march, wait, peer composite, signal) for A/B")
    ap.add_argument("--balance", type=int, default=1,
                    help="sort-last: 1 = move the slab cuts to equal MEASURED march time at set-up (ownership shift inside "
                         "the slabs' margins, no data movement); 0 = keep the analytic view-balanced cuts")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU-baseline band (0 = auto)")
    ap.add_argument("--extra", type=int, default=1, help="also measure the secondary variants (N=1 only)")
    ap.add_argument("--nvdb-codec", default="float", choices=["float", "fp4", "fp8", "fp16", "fpn"],
                    help="c5 only: NanoVDB grid type of the fog sphere (quantised by visrtx_b200.nvdb_writer)")
    ap.add_argument("--c4-scaling", type=int, default=-1,
                    help="N = 8 runs: also measure BASELINE config C4 (4096^3 f32, 256 GiB, sort-last) on 2, 4 and 8 of the "
                         "ranks (extra.c4_scaling); -1 = on for the default C2 run at N = 8, 0 = off")
    ap.add_argument("--c3-sort-first", type=int, default=-1,
                    help="after the headline run on N > 1 GPUs, also render BASELINE config C3 (2048^3 UFIXED16, 4K) "
                         "sort-first on all ranks (extra.c3_sort_first); -1 = on for the default C2 sort-last run, 0 = off")
    ap.add_argument("--anari-multi-gpu", type=int, default=-1,
                    help="N > 1: rank 0 also drives ALL N GPUs through the ANARI C API in one process (device parameter "
                         "cudaDevices, sort-last) while the other ranks wait (extra.anari_multi_gpu); -1 = on for the "
                         "default C2 sort-last run, 0 = off")
    ap.add_argument("--save-frame", default="", help="rank 0: write frame 0 of the timed scene (uint32 sRGB8) to this .npy")
    ap.add_argument("--extra-configs", default="c3,c5",
                    help="N=1 default run only: other BASELINE configs measured after the headline (extra.configs), "
                         "each with its own roofline and parity block; empty string = none")
    a = ap.parse_args()
    return apply_preset(a)


def apply_preset(a):
    if a.config == "c3":
        a.size, a.width, a.height, a.field = 2048, 3840, 2160, "shells"
        a.skip = 1 if a.skip < 0 else a.skip
        a.unit_distance = 8.0
    elif a.config == "c4":
        a.size, a.unit_distance, a.mode = 4096, 1024.0, "sort-last"
    elif a.config == "c5":  # NanoVDB fog sphere r=200 voxels, 1080p, progressive accumulation
        a.field, a.size, a.unit_distance = "fog", 401, 64.0
        a.skip = 1 if a.skip < 0 else a.skip
    a.skip = max(a.skip, 0)
    return a


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
def make_scene(args, torch, device, z_begin=0, z_end=None):
    """Analytic field on a size^3 lattice, origin 0, spacing 1, generated in HBM.
    ml: Marschner-Lobb f32 (dense).  shells: 12 thin Gaussian shells, UFIXED16 (sparse, C3)."""
    from visrtx_b200 import scenes
    n = args.size
    if args.field == "shells":
        assert z_begin == 0 and z_end is None, "the sparse field is generated whole"
        return scenes.shells_torch(n, device)
    if args.field == "fog":  # a serialized NanoVDB float grid (host numpy uint8), written by our own writer
        from visrtx_b200 import nvdb_writer
        return nvdb_writer.fog_sphere(radius=(n - 1) / 2.0, voxel_size=1.0, half_width=3.0, codec=args.nvdb_codec)
    return scenes.marschner_lobb_torch(n, device, z_begin=z_begin, z_end=z_end, nz_total=n)


def scene_dtype(args):
    from visrtx_b200 import pods
    return pods.DVR_UFIXED16 if args.field == "shells" else pods.DVR_FLOAT32


def voxel_bytes(args):
    return 2 if args.field == "shells" else 4


def scene_bounds(args):
    """object-space bounds of the field (what dvr_field_bounds returns)."""
    n = args.size
    if args.field == "fog":
        r = (n - 1) // 2
        return (-float(r) + 1.0,) * 3, (float(r),) * 3  # world bbox of the active voxels [-199, 200) for r=200
    return (0.0, 0.0, 0.0), (n - 1.0,) * 3


def create_field(args, capi, vol, stream):
    n = args.size
    if args.field == "fog":
        return capi.Field.create_nanovdb(vol.ctypes.data, vol.nbytes, False, stream)
    return capi.Field.create_structured(vol.data_ptr(), True, scene_dtype(args), (n, n, n), (0, 0, 0), (1, 1, 1),
                                        capi.DVR_FILTER_LINEAR, stream)


def scene_colormap(args):
    from visrtx_b200 import scenes
    return scenes.sparse_colormap(256, 0.5) if args.field == "shells" else scenes.tsd_default_colormap(256)


def workload_name(args):
    n, W, H = args.size, args.width, args.height
    field = ("UFIXED16 sparse shells (12 Gaussian shells r=96*n/2048 voxels, zero elsewhere)" if args.field == "shells"
             else "f32 Marschner-Lobb (dense)")
    if args.field == "fog":
        return (f"{args.config.upper()}: NanoVDB {args.nvdb_codec} fog sphere r={(n - 1) // 2} voxels (index bbox {n}^3, 33.5 M active "
                f"voxels at r=200) + transferFunction1D (TSD default map, unitDistance {args.unit_distance:g}), {W}x{H}, "
                f"default renderer 1 spp progressive accumulation, volumeSamplingRate {args.rate:g}, orbit camera "
                f"az30/el20 at 2|diag|, fovy 60")
    tfn = "sparse map (alpha 0 below 0.5)" if args.field == "shells" else "TSD default map"
    return (f"{args.config.upper()}: {n}^3 {field} structuredRegular + transferFunction1D ({tfn}, unitDistance "
            f"{args.unit_distance:g} voxels), {W}x{H}, default renderer 1 spp progressive, volumeSamplingRate "
            f"{args.rate:g} (step {0.5 / args.rate:g} voxel), orbit camera az30/el20 at 2|diag|, fovy 60")


def orbit(args, az_deg=30.0, oracle_side=False):
    """Benchmark camera.  oracle_side=True builds the POD with the oracle's camera set-up (bit-identical to
    dvr_camera_perspective, tests/test_capi_host.py) so that the reference arm never loads libdvr_b200.so."""
    from visrtx_b200 import scenes
    lo, hi = scene_bounds(args)
    pose = scenes.orbit_camera(lo, hi, args.width, args.height, az_deg=az_deg)
    if oracle_side:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob
        return ob.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect), pose
    from visrtx_b200 import capi
    return capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect), pose


def oracle_tf(args):
    """The 256-texel table from the REFERENCE's own discretisation (libref_host.so), else O-cpu's restatement of it —
    both bit-identical to dvr_tf_discretize (tests/test_capi_host.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    return ob.tf_discretize(color=scene_colormap(args), which="ref" if ob.have_ref_host() else "cpu")


def parity_block(torch, fmt_bytes, color_a, color_b, depth_a=None, depth_b=None, what="", early_terminated=None):
    """north_star's tolerance on one frame: per-pixel max |d| in 1/255 units after tonemap/encode, PSNR; depth."""
    a = color_a.view(torch.uint8).view(-1, 4).to(torch.int32)
    b = color_b.view(torch.uint8).view(-1, 4).to(torch.int32)
    d = (a - b).abs()
    dmax = d.max(dim=1).values
    mse = float(((d.to(torch.float64) / 255.0) ** 2).mean().item())
    out = {"max_abs_255": int(dmax.max().item()), "frac_le_2": float((dmax <= 2).double().mean().item()),
           "frac_identical": float((dmax == 0).double().mean().item()),
           "psnr_db": 99.0 if mse == 0.0 else 10.0 * math.log10(1.0 / mse),
           "pixels_gt_2": int((dmax > 2).sum().item()), "tolerance": "north_star: <= 2/255 per pixel, PSNR >= 45 dB",
           "what": what}
    if depth_a is not None and depth_b is not None:
        hit = (depth_a < 1e29) & (depth_b < 1e29)
        rel = ((depth_a - depth_b).abs() / depth_b.abs().clamp_min(1e-20))[hit]
        out["depth_max_rel"] = float(rel.max().item()) if rel.numel() else 0.0
        out["depth_hit_mask_equal"] = bool(((depth_a < 1e29) == (depth_b < 1e29)).all().item())
    out["pass"] = bool(out["max_abs_255"] <= 2 and out["psnr_db"] >= 45.0)
    if early_terminated is not None:
        # Sort-last only.  A ray whose one-pass march stops at opacity >= 0.99 INSIDE a slab (gpu/volumeIntegration.h:86)
        # keeps, in the slab decomposition, that slab's samples behind the stopping point: the slab starts from opacity 0
        # and cannot know what lies in front of it.  Their weight is <= 1 - 0.99, i.e. <= 0.01 x colour in linear space
        # (<= 2.55/255 before the sRGB curve, a few steps after it on dark pixels).  Every other pixel must meet
        # north_star's bound; the early-terminated ones are counted and bounded separately.
        et = early_terminated.view(-1)
        out["early_terminated_rays"] = int(et.sum().item())
        out["pixels_gt_2_not_early_terminated"] = int(((dmax > 2) & ~et).sum().item())
        out["max_abs_255_not_early_terminated"] = int(dmax[~et].max().item()) if bool((~et).any().item()) else 0
        out["tolerance"] = ("north_star (<= 2/255 per pixel, PSNR >= 45 dB) on every ray the one-pass march does not "
                            "terminate early; <= 6/255 on early-terminated rays (sort-last keeps <= 0.01 x colour of "
                            "samples behind the stopping point, see DESIGN.md 6)")
        out["pass"] = bool(out["pixels_gt_2_not_early_terminated"] == 0 and out["max_abs_255"] <= 6
                           and out["psnr_db"] >= 45.0)
    return out


def bytes_per_frame(args, cells_touched: int, samples: int = 0, fmt_bytes: int = 4):
    """N_vox_touched * sizeof(voxel) + N_px * (32 accum RW + colour + 8 depth RW) + 4 KiB TF, with
    N_vox_touched = voxels of the macrocells the rays fetched from.  When rays are many voxels apart (C4 at
    1080p: 15 voxels) that over-counts: a sample cannot pull more than 4 32-byte sectors, so the voxel term is
    capped at samples * 128 B.  Returns (bytes, which_model)."""
    frame = args.width * args.height * (32 + fmt_bytes + 8) + 4096
    cell_model = cells_touched * 16 ** 3 * voxel_bytes(args)
    if samples and samples * 128 < cell_model:
        return samples * 128 + frame, "sector cap: samples*128B (rays sparser than macrocells)"
    return cell_model + frame, "macrocells touched * 16^3 * sizeof(voxel) (SURVEY 8d)"


def traffic_key(args, mode):
    codec = f":{args.nvdb_codec}" if args.field == "fog" and args.nvdb_codec != "float" else ""
    return (f"{args.config}:{args.field}{codec}:{args.size}:{args.width}x{args.height}:rate{args.rate:g}:"
            f"ud{args.unit_distance:g}:skip{int(bool(args.skip))}:{mode}")


def measured_traffic(args, mode):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this exact
    workload (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(traffic_key(args, mode))
        return None if e is None else int(e["dram_bytes_per_launch"])
    except (OSError, ValueError, KeyError):
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
def cpu_baseline(args, torch, vol_dev, samples_per_frame, rows=0):
    """O-cpu on a bounded band of rows around the image centre (all host cores via OpenMP).  samples_per_frame None:
    frames/s is scaled by rows instead of by samples (no GPU-side sample count available)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from visrtx_b200 import pods as capi
    n = args.size
    host = vol_dev if args.field == "fog" else vol_dev.cpu().numpy()  # (z,y,x) voxels, or the NanoVDB blob
    if args.field == "shells":  # UFIXED16 stored as int16 bits -> what cudaReadModeNormalizedFloat hands the filter
        host = (host.view(np.uint16).astype(np.float32) / np.float32(65535.0))
    tf = oracle_tf(args)
    cam, _ = orbit(args, oracle_side=True)
    vols = (ob.OracleVolume * 1)()
    o = vols[0]
    o.voxels = host.ctypes.data_as(C.c_void_p)
    if args.field == "fog":
        o.nvdbGrid = host.ctypes.data_as(C.c_void_p)
    o.dims = (C.c_int32 * 3)(n, n, n)
    o.origin = (C.c_float * 3)(0, 0, 0)
    o.spacing = (C.c_float * 3)(1, 1, 1)
    o.tf = tf.ctypes.data_as(C.c_void_p)
    o.valueRange = (C.c_float * 2)(0, 1)
    o.unitDistance = args.unit_distance
    o.id = 0
    o.worldToObject = (C.c_float * 12)(*capi.IDENTITY_3X4)
    o.instanceId = 0
    npx = args.width * args.height
    accum = np.zeros((npx, 4), np.float32)
    color = np.zeros(npx, np.uint32)
    depth = np.zeros(npx, np.float32)
    b = ob.OracleBuffers()
    b.colorAccumulation = accum.ctypes.data_as(C.c_void_p)
    b.outColor = color.ctypes.data_as(C.c_void_p)
    b.depth = depth.ctypes.data_as(C.c_void_p)
    p = capi.frame_params(args.width, args.height, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, 0,
                          -1, 1, args.rate, (0.1, 0.1, 0.1, 1.0))
    cores = os.cpu_count() or 1
    rows = rows or args.cpu_rows
    mid = args.height // 2
    lib = ob.cpu()

    def run(nrows):
        s = C.c_uint64()
        t0 = time.perf_counter()
        lib.oracle_render(C.byref(p), C.byref(cam), vols, 1, C.byref(b), C.byref(s), mid - nrows // 2,
                          mid - nrows // 2 + nrows)
        return time.perf_counter() - t0, s.value

    if rows <= 0:  # calibrate on 2 rows (also warms the page cache of the host volume), then size the band
        run(2)
        dt, _ = run(2)
        rows = int(max(2, min(args.height, 2 * 15.0 / max(dt, 1e-3))))
    # repeat the band until ~12 s of CPU work have been timed (bounded sample)
    total_dt, total_samp, reps = 0.0, 0, 0
    while total_dt < 12.0 and reps < 200:
        dt, nsamp = run(rows)
        total_dt += dt
        total_samp += nsamp
        reps += 1
    dt, nsamp = total_dt, total_samp
    sps = nsamp / dt
    value = (sps / max(samples_per_frame, 1)) if samples_per_frame else (rows * reps / dt) / args.height
    return {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
            "gsamples_per_s": sps / 1e9,
            "sample": f"{rows} image rows around the centre of the same {args.width}x{args.height} frame, repeated "
                      f"{reps}x ({nsamp} samples, {dt:.1f} s of CPU work); frames/s = CPU samples/s / samples per frame"}


# ---------------------------------------------------------------------------------------------------------
class AnariE2E:
    """The C2 scene built through the ANARI C API (what an application does), for the e2e number."""

    def __init__(self, args, torch, device, vol_dev, rank, world, mode, gpus=None):
        from visrtx_b200 import anari as A
        from visrtx_b200 import scenes
        self.A, self.args, self.torch = A, args, torch
        n, W, H = args.size, args.width, args.height
        d = self.d = A.Device()
        if gpus:  # one process, several GPUs behind the same anari* calls (display GPU first)
            d.set(d.handle, "cudaDevices", A.STRING, ",".join(str(g) for g in gpus))
            d.set(d.handle, "multiGpuMode", A.STRING, "sortLast")
        else:
            d.set(d.handle, "cudaDevice", A.INT32, device.index)
        d.commit(d.handle)
        torch.cuda.synchronize()
        if args.field == "fog":
            self.data = d.new_array1d(vol_dev, A.UINT8)
            self.field = d.new("SpatialField", "nanovdb")
            d.set(self.field, "data", A.ARRAY1D, self.data)
        else:
            self.data = d.new_array3d_device(vol_dev.data_ptr(), A.UFIXED16 if args.field == "shells" else A.FLOAT32,
                                             n, n, n)
            self.field = d.new("SpatialField", "structuredRegular")
            d.set(self.field, "data", A.ARRAY3D, self.data)
        d.commit(self.field)
        self.volume = d.new("Volume", "transferFunction1D")
        self.color = d.new_array1d(scene_colormap(args), A.FLOAT32_VEC4)
        d.set(self.volume, "color", A.ARRAY1D, self.color)
        d.set(self.volume, "value", A.SPATIAL_FIELD, self.field)
        d.set(self.volume, "unitDistance", A.FLOAT32, args.unit_distance)
        d.commit(self.volume)
        self.world = d.new("World")
        self.vols = d.new_object_array([self.volume], A.VOLUME)
        d.set(self.world, "volume", A.ARRAY1D, self.vols)
        d.commit(self.world)
        self.camera = d.new("Camera", "perspective")
        self.renderer = d.new("Renderer", "default")
        d.set(self.renderer, "background", A.FLOAT32_VEC4, (0.1, 0.1, 0.1, 1.0))
        d.set(self.renderer, "volumeSamplingRate", A.FLOAT32, args.rate)
        d.set(self.renderer, "sampleLimit", A.INT32, 0)
        d.set(self.renderer, "macrocellSkipping", A.BOOL, int(bool(args.skip)))
        if mode == "sort-first":
            d.set(self.renderer, "sortFirstRank", A.INT32, rank)
            d.set(self.renderer, "sortFirstRanks", A.INT32, world)
        d.commit(self.renderer)
        self.frame = d.new("Frame")
        d.set(self.frame, "size", A.UINT32_VEC2, (W, H))
        d.set(self.frame, "channel.color", A.DATA_TYPE, A.UFIXED8_RGBA_SRGB)
        d.set(self.frame, "channel.depth", A.DATA_TYPE, A.FLOAT32)
        d.set(self.frame, "renderer", A.RENDERER, self.renderer)
        d.set(self.frame, "camera", A.CAMERA, self.camera)
        d.set(self.frame, "world", A.WORLD, self.world)
        d.commit(self.frame)
        self.w, self.h, self.t = C.c_uint32(), C.c_uint32(), C.c_int()
        self.checksum = 0

    def prepare(self, n_steps):
        """The per-step inputs (camera poses of the orbit) as ready-made parameter buffers: generating the camera
        path is the application's business, not part of the measured API calls."""
        A, args = self.A, self.args
        self.inputs = []
        for i in range(n_steps):
            _, pose = orbit(args, az_deg=30.0 + 0.05 * i)
            bufs = []
            for name, dt, val in ((b"position", A.FLOAT32_VEC3, pose.position), (b"direction", A.FLOAT32_VEC3, pose.direction),
                                  (b"up", A.FLOAT32_VEC3, pose.up), (b"fovy", A.FLOAT32, (pose.fovy,)),
                                  (b"aspect", A.FLOAT32, (pose.aspect,))):
                arr = (C.c_float * len(val))(*[float(v) for v in val])
                bufs.append((name, dt, arr, C.cast(arr, C.c_void_p)))
            self.inputs.append(bufs)

    def step(self, i):
        A, d = self.A, self.d
        for name, dt, _keep, ptr in self.inputs[i % len(self.inputs)]:
            A.lib.anariSetParameter(d.handle, self.camera, name, dt, ptr)
        A.lib.anariCommitParameters(d.handle, self.camera)
        A.lib.anariRenderFrame(d.handle, self.frame)
        A.lib.anariFrameReady(d.handle, self.frame, A.WAIT)
        p = A.lib.anariMapFrame(d.handle, self.frame, b"channel.color", C.byref(self.w), C.byref(self.h), C.byref(self.t))
        # read the result on the host: one pixel per step is enough to prove the bytes arrived
        self.checksum ^= C.cast(p, C.POINTER(C.c_uint32))[(self.w.value * self.h.value) // 2]
        A.lib.anariUnmapFrame(d.handle, self.frame, b"channel.color")

    def bytes_per_step(self):
        # host -> device: camera parameters (5 setParameter calls) + the kernel parameter block; device -> host: colour
        return 3 * 12 + 2 * 4 + 2048, self.args.width * self.args.height * 4

    def close(self):
        d = self.d
        errs = [m for m in d.messages if m[0] <= self.A.SEVERITY_ERROR]
        if errs:
            raise RuntimeError(f"ANARI device reported errors: {errs[:3]}")
        for o in (self.frame, self.renderer, self.camera, self.world, self.vols, self.volume, self.color, self.field,
                  self.data):
            d.release(o)
        d.close()


# ---------------------------------------------------------------------------------------------------------
def run_ours(args, torch, dist, rank, world):
    from visrtx_b200 import capi, scenes
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(device)
    capi.set_device(device.index)
    stream = torch.cuda.current_stream().cuda_stream
    n, W, H = args.size, args.width, args.height
    npx = W * H

    mode = args.mode
    if mode == "auto":
        # sort-last (z-slabs) also wins when the volume fits one GPU: it shortens every ray by 1/N, while
        # sort-first leaves the longest ray (the kernel's critical path) untouched (profiles/r01_multigpu.md)
        mode = "sort-last" if world > 1 else "single"
    if world == 1 and mode == "sort-first":
        mode = "single"

    t_setup = time.perf_counter()
    field_commit_ms = None
    vol = None
    if mode == "sort-last":
        from visrtx_b200 import multigpu as _mg
        # slabs of equal WORK for the benchmark camera (sample density falls off as 1/r^2 from the eye), not of
        # equal thickness: see multigpu.view_balanced_slab_ranges
        _lo, _hi = scene_bounds(args)
        _, _pose0 = orbit(args)
        slab_ranges0 = _mg.view_balanced_slab_ranges(n, world, _lo, _hi, _pose0.position)
        # every slab keeps a margin of slices resident beyond its initial range: ownership can then move between
        # neighbours without moving voxels (feedback balancing below)
        slab_margin = _mg.slab_margin(n, world) if (world > 1 and args.balance) else 0
        slab_limits = _mg.creation_ranges(slab_ranges0, n, slab_margin)
        z0, z1 = slab_limits[rank]
        r0, r1 = _mg.resident_range(z0, z1, n)
        field = capi.Field.create_slab(0, True, scene_dtype(args), (n, n, n), z0, z1, (0, 0, 0), (1, 1, 1),
                                       capi.DVR_FILTER_LINEAR, stream)
        field.set_owned_slices(*slab_ranges0[rank])
        chunk = 32
        for zc in range(r0, r1, chunk):  # generate + upload in chunks: the slab is never staged twice in HBM
            ze = min(zc + chunk, r1)
            part = make_scene(args, torch, device, z_begin=zc, z_end=ze)
            field.upload_slices(part.data_ptr(), True, zc - r0, ze - zc, stream)
            torch.cuda.synchronize()
            del part
        field.build_macrocells(stream)
    else:
        vol = make_scene(args, torch, device)
        torch.cuda.synchronize()
        t_commit = time.perf_counter()
        field = create_field(args, capi, vol, stream)
        torch.cuda.synchronize()
        # what a time-varying field pays per update: array allocation + copy into it + macrocell ranges (K4)
        field_commit_ms = (time.perf_counter() - t_commit) * 1e3
    torch.cuda.synchronize()
    tf = capi.tf_discretize(color=scene_colormap(args))
    volume = capi.Volume.create(field, tf, (0.0, 1.0), args.unit_distance, 0, stream)
    inst, ninst = capi.make_instances([volume], None, [0])
    cam, _ = orbit(args)
    setup_s = time.perf_counter() - t_setup

    from visrtx_b200 import multigpu
    FMT, INTEG, BG = capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, (0.1, 0.1, 0.1, 1.0)
    if mode == "sort-last":
        driver = multigpu.SortLast(capi, torch, dist, rank, world, device, W, H, inst, 0, 0, FMT, INTEG, args.rate, BG,
                                   skip=bool(args.skip), host_mirror=world > 1, fused=bool(args.fused))
    else:
        driver = multigpu.SortFirst(capi, torch, dist, rank, world, device, W, H, inst, ninst, FMT, INTEG, args.rate,
                                    BG, skip=bool(args.skip), tile_band=int(os.environ.get("DVR_TILE_BAND", "1")),
                                    host_mirror=world > 1)
    driver.stream_to_host(False)  # the device-timed region keeps every byte in HBM; e2e turns the host stream on
    slab_balance = None
    if mode == "sort-last" and world > 1:
        slab_balance = {"margin_slices": slab_margin, "initial": [list(r) for r in slab_ranges0],
                        "what": "initial cuts: equal 1/r^2 weight for the camera (view_balanced_slab_ranges)"}
        if args.balance:
            final, times, hist = driver.calibrate(field, slab_ranges0, slab_limits, cam, stream)
            slab_balance.update({"final": [list(r) for r in final], "rounds": hist,
                                 "what": "cuts moved to equal MEASURED march time per GPU (SortLast.calibrate: "
                                         "dvr_render_partial timed per rank, all-gathered, ownership shifted inside the "
                                         "slabs' resident margins with dvr_field_set_owned_slices; no voxel moves)"})
    fb = driver.fb
    host_color = torch.empty(npx, dtype=torch.int32, pin_memory=True)

    def params(frame_id):
        return driver.params(frame_id)

    def step(frame_id):
        driver.render(frame_id, cam, stream)

    # ---- untimed instrumented launch: samples / touched macrocells of this workload
    stats_t = torch.zeros(4, dtype=torch.int64, device=device)
    p_stats = params(0)
    p_stats.tileRank, p_stats.tileRanks = 0, 1
    scratch_color = torch.zeros(npx, dtype=torch.int32, device=device)
    fb_stats = capi.frame_buffers(driver.accum.data_ptr(), scratch_color.data_ptr(), driver.depth.data_ptr())
    if mode == "sort-last":
        capi.render_partial_instrumented(p_stats, cam, inst, driver.rgba_ptrs[0][rank], driver.depth_ptrs[0][rank],
                                         stats_t.data_ptr(), stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(stats_t)  # samples / touched cells summed over the slabs
    else:
        capi.render_instrumented(p_stats, cam, inst, ninst, fb_stats, stats_t.data_ptr(), stream)
        torch.cuda.synchronize()
    samples, skipped, rays_hit, cells = stats_t.tolist()

    # ---- device-timed region (clocks are sampled from here to the end of the e2e loop: the timed
    # region alone is ~0.1 s, shorter than nvidia-smi's sampling period)
    sampler = ClockSampler(device.index)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = capi.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    fps = 1000.0 / ms_per_step

    # ---- kernel-only duration for the roofline (render launches only, no exchange)
    if mode == "sort-last":
        krender = lambda fid: capi.render_partial(params(fid), cam, inst, driver.rgba_ptrs[0][rank],
                                                  driver.depth_ptrs[0][rank], stream)
    else:
        krender = lambda fid: capi.render(params(fid), cam, inst, ninst, fb_stats, stream)
    for i in range(3):
        krender(1 + i)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kn = max(10, min(args.steps, 50))
    k0.record()
    for i in range(kn):
        krender(10 + i)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / kn
    kernel_ms_per_rank = [kernel_ms]
    if world > 1:
        t = torch.tensor([kernel_ms], dtype=torch.float64, device=device)
        allk = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allk, t)
        kernel_ms_per_rank = [float(v.item()) for v in allk]
        kernel_ms = max(kernel_ms_per_rank)

    # ---- per-frame distribution (SURVEY 8d timing protocol): progressive accumulation vs single-shot (reset each
    # frame), one CUDA-event pair per frame, median and p95
    frame_dist = None
    if world == 1 and args.extra:
        frame_dist = {}
        for label, fid_of in (("progressive", lambda i: 100 + i), ("single_shot", lambda i: 0)):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
            for i in range(5):
                step(fid_of(i))
            torch.cuda.synchronize()
            for i, (a_, b_) in enumerate(evs):
                a_.record()
                step(fid_of(i))
                b_.record()
            torch.cuda.synchronize()
            ts = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
            frame_dist[label] = {"median_ms": ts[len(ts) // 2], "p95_ms": ts[int(len(ts) * 0.95)], "min_ms": ts[0],
                                 "frames": len(ts)}

    # ---- end-to-end with host buffers.  N=1: through the ANARI C API of the device library (what an
    # application calls).  N>1: the multi-GPU driver (C-ABI) with a moved camera every step and the
    # display rank mapping the assembled frame to pinned host memory.
    import ctypes as _C
    _cudart = _C.CDLL("libcudart.so")
    if world == 1 and mode == "single":
        e2e = AnariE2E(args, torch, device, vol, rank, world, mode)
        e2e_step = e2e.step
        e2e_what = ("ANARI C API of libanari_library_visrtx_b200.so: anariSetParameter(camera)+anariCommitParameters"
                    "+anariRenderFrame+anariFrameReady(WAIT)+anariMapFrame(channel.color -> host) per step, wall "
                    "clock; the volume is an ANARI_NV_ARRAY_CUDA shared array uploaded once; after the first host map "
                    "the device streams the encoded colour into pinned host memory during the launch "
                    "(DvrFrameBuffers::outColorMirror), so the device->host bytes overlap the march")
    else:
        e2e = None

        e2e_cams = [orbit(args, az_deg=30.0 + 0.05 * i)[0] for i in range(args.steps + 3)]

        def e2e_step(i):
            cam_i = e2e_cams[i % len(e2e_cams)]
            driver.render(0, cam_i, stream)
            if rank == 0 and driver.host_frame is None:
                _cudart.cudaMemcpyAsync(_C.c_void_p(host_color.data_ptr()), _C.c_void_p(driver.color_ptr),
                                        _C.c_size_t(npx * 4), _C.c_int(2), _C.c_void_p(stream))
            torch.cuda.synchronize()
            if rank == 0 and driver.host_frame is not None:  # read the result on the host
                e2e_state["checksum"] ^= int(host_view[npx // 2])

        e2e_state = {"checksum": 0}
        driver.stream_to_host(True)
        host_view = driver.host_frame.numpy() if driver.host_frame is not None else None
        e2e_what = (f"{mode} driver over the C-ABI: moved camera (kernel parameter upload) + render on {world} GPUs; "
                    "every rank's final-colour stores also go to one pinned host frame shared by all processes "
                    "(POSIX shm + cudaHostRegister, DvrFrameBuffers::outColorMirror), read on the display rank "
                    "every step, wall clock" if driver.host_frame is not None else
                    f"{mode} driver over the C-ABI: moved camera + render on {world} GPUs + assembled colour frame "
                    "copied to pinned host memory on the display rank every step, wall clock")

    if e2e is not None:
        e2e.prepare(args.steps + 3)
    for i in range(3):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ne2e = args.steps
    for i in range(ne2e):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_bytes = e2e.bytes_per_step() if e2e is not None else (2048 * world, npx * 4)
    if e2e is not None:
        e2e.close()
    e2e_fps = ne2e / e2e_s
    clocks = sampler.stop() if rank == 0 else None

    peak, peak_src = measured_peak()
    bframe, bmodel = bytes_per_frame(args, cells, samples)
    share = 1.0 / world
    # per-GPU achieved bandwidth: every rank reads ~1/N of the touched voxels (its tile rows / its slab)
    achieved = bframe * share / (kernel_ms * 1e-3) / 1e9
    out = {
        "metric": metric_name(args),
        "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (analytic field generated in HBM: " + args.field + ")",
        "config": bench_config(args, mode, world),
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": e2e_bytes[0],
                "d2h_bytes_per_step": e2e_bytes[1],
                "what": e2e_what},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(args, mode), "peak_source": peak_src, "kernel": "dvrFrameKernel",
                     "kernel_ms": kernel_ms, "algorithmic_bytes": bframe, "bytes_model": bmodel,
                     "macrocells_touched": int(cells), "macrocells_total": int(math.ceil(n / 16) ** 3)},
        "extra": {"samples_per_frame": int(samples), "gsamples_per_s": samples * fps / 1e9 * (1 if world == 1 else 1),
                  "rays_hit": int(rays_hit), "bytes_per_sample": bframe / max(samples, 1), "setup_s": setup_s,
                  "share_per_rank": share},
    }
    if mode == "sort-last" and world > 1 and args.fused:
        # phases of the fused launch on every rank (%globaltimer stamps of 20 extra, untimed frames), in the
        # configuration of the device-timed region: the e2e loop left the host mirror on, and its PCIe stores stretch
        # the march phase rank by rank (until call AB the phases were stamped with it: the display rank looked 40-80 us
        # slower than the others, profiles/r02_sort_last_fused.md)
        driver.stream_to_host(False)
        nt = 20
        tbuf = torch.zeros((nt, 8), dtype=torch.int64, device=device)
        tbuf[:, 0] = torch.iinfo(torch.int64).max
        for i in range(nt):
            driver.timing_ptr = tbuf[i].data_ptr()
            driver.render(200 + i, cam, stream)
        driver.timing_ptr = 0
        torch.cuda.synchronize()
        dist.barrier()
        t = tbuf.double()
        ph = torch.stack([(t[:, 1] - t[:, 0]), (t[:, 2] - t[:, 1]), (t[:, 3] - t[:, 2]), (t[:, 4] - t[:, 3]),
                          (t[:, 4] - t[:, 0]), (t[:, 6] - t[:, 0]), (t[:, 5] - t[:, 0])], dim=1).median(dim=0).values / 1e3  # us
        allph = [torch.zeros_like(ph) for _ in range(world)]
        dist.all_gather(allph, ph)
        out["extra"]["fused_phases_us_per_rank"] = {
            "columns": ["march (first CTA -> last tile)", "background strip", "owned regions composited (incl. waiting "
                        "for the slowest rank's flags)", "retire (display rank: all ranks' flags)", "kernel total", "own last region flag published (since start)",
                        "last owned region seen complete on all ranks (since start)"],
            "ranks": [[round(float(v), 1) for v in a.tolist()] for a in allph]}
    if mode == "sort-last" and world > 1:
        # the slowest rank's march alone (dvr_render_partial, no exchange) against the whole step: what is left is
        # exchange + compositing + imbalance that the fused launch could not hide
        out["extra"]["slab_balance"] = slab_balance
        out["extra"]["march_us"] = kernel_ms * 1e3
        out["extra"]["march_alone_us_per_rank"] = [round(v * 1e3, 1) for v in kernel_ms_per_rank]
        out["extra"]["exchange_us"] = (ms_per_step - kernel_ms) * 1e3
        out["extra"]["sort_last_launch"] = ("fused: dvr_render_slab_frame, one launch per GPU and frame" if args.fused else
                                            "legacy: partial march + wait + peer composite + signal")
        out["extra"]["slabs"] = ("z-slabs of equal work for the initial camera (sample density ~ 1/r^2 from the eye), not "
                                 "of equal thickness (multigpu.view_balanced_slab_ranges)")
    if clocks is not None:
        out["clocks"] = clocks
    if frame_dist is not None:
        out["extra"]["frame_time_ms"] = frame_dist
    if field_commit_ms is not None:
        out["extra"]["field_commit_ms"] = field_commit_ms

    if rank == 0 and world == 1 and mode == "single" and args.extra:
        out["extra"]["variants"] = measure_variants(args, torch, capi, scenes, field, cam, inst, ninst, fb, stream, stats_t)
    if rank == 0 and world == 1 and vol is not None and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_baseline(args, torch, vol, samples)
        except Exception as e:  # the oracle is optional at bench time
            out["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"unavailable: {e}"}
    if rank == 0 and world == 1 and vol is not None:
        # same-config parity: frame 0 of the benchmark camera rendered by both arms on this very scene
        try:
            (ref_color, ref_depth), ref_fps = ref_gpu_frame0(args, torch, vol, min(args.steps, 30) if args.extra else 0)
            par_color = torch.zeros(npx, dtype=torch.int32, device=device)
            par_depth = torch.zeros(npx, dtype=torch.float32, device=device)
            fb_par = capi.frame_buffers(driver.accum.data_ptr(), par_color.data_ptr(), par_depth.data_ptr())
            capi.render(params(0), cam, inst, ninst, fb_par, stream)
            torch.cuda.synchronize()
            out["parity"] = parity_block(torch, 4, par_color, ref_color, par_depth, ref_depth,
                                         what="frame 0 of the timed scene and camera: this arm (dvr_render) vs O-gpu "
                                              "(the reference's device code) on the same B200, sRGB8 colour + depth")
            if ref_fps is not None:
                out["extra"]["ref_gpu_fps"] = ref_fps
            del ref_color, ref_depth, par_color, par_depth
        except Exception as e:
            out["parity"] = {"pass": None, "what": f"unavailable: {e}"}
    if rank == 0 and world == 1 and vol is not None and args.extra:
        if mode == "single":
            try:
                out["extra"]["dpt"] = measure_dpt(args, torch, capi, field, cam, fb, stream, vol)
            except Exception as e:
                out["extra"]["dpt"] = f"unavailable: {e}"
            if args.field == "ml":
                try:
                    out["extra"]["time_varying"] = measure_time_varying(args, torch, capi, field, cam, fb, stream, vol)
                except Exception as e:
                    out["extra"]["time_varying"] = f"unavailable: {e}"
    if rank == 0 and world == 1 and mode == "single" and args.extra and args.config == "c2" and args.extra_configs:
        # the other single-GPU BASELINE configs, driver-observed: free the headline scene first
        volume.destroy()
        field.destroy()
        del vol
        torch.cuda.empty_cache()
        out["extra"]["configs"] = {}
        for name in [c for c in args.extra_configs.split(",") if c in ("c3", "c5")]:
            try:
                out["extra"]["configs"][name] = measure_secondary_config(args, name, torch, device, stream)
            except Exception as e:
                out["extra"]["configs"][name] = f"unavailable: {type(e).__name__}: {e}"
    if rank == 0 and args.save_frame:
        driver.stream_to_host(False)
    if args.save_frame:
        driver.render(0, cam, stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if rank == 0:
            np.save(args.save_frame, driver.color_tensor().cpu().numpy().view(np.uint32))
    if world > 1:
        out["parity_vs_single"] = parity_vs_single(args, torch, dist, capi, driver, cam, rank, world, device, stream, mode)
    if driver is not None and getattr(driver, "host_frame", None) is not None:
        host_view = None  # drop the numpy view before the shared segment is unmapped
        driver.host_frame.close(dist if world > 1 else None)
        driver.host_frame = None
    do_c4 = args.c4_scaling == 1 or (args.c4_scaling < 0 and world == 8 and args.config == "c2" and mode == "sort-last")
    do_c3 = args.c3_sort_first == 1 or (args.c3_sort_first < 0 and world >= 2 and args.config == "c2" and mode == "sort-last")
    if (do_c4 or do_c3) and world >= 2:
        # free the headline scene on every rank first: a C4 slab is up to 137 GB per GPU, a C3 replica 16 GiB (+ 16 GiB
        # of staging while it is generated)
        driver.close()
        volume.destroy()
        field.destroy()
        torch.cuda.empty_cache()
    if do_c3 and world >= 2:
        try:
            c3 = measure_c3_sort_first(args, torch, dist, capi, rank, world, device, stream)
        except Exception as e:  # a secondary measurement must never take the headline line down
            c3 = f"unavailable: {type(e).__name__}: {e}"
        if rank == 0:
            out["extra"]["c3_sort_first"] = c3
    do_anari = world >= 2 and (args.anari_multi_gpu == 1 or (args.anari_multi_gpu < 0 and args.config == "c2"
                                                             and mode == "sort-last"))
    if do_anari and not (do_c4 or do_c3):
        driver.close()
        volume.destroy()
        field.destroy()
        torch.cuda.empty_cache()
    if do_anari:
        am = measure_anari_multi_gpu(args, torch, dist, rank, world, device)
        if rank == 0:
            out["extra"]["anari_multi_gpu"] = am
    if do_c4 and world >= 2:
        c4 = measure_c4_scaling(args, torch, dist, capi, rank, world, device, stream)
        if rank == 0:
            out["extra"]["c4_scaling"] = c4
    return out


def measure_anari_multi_gpu(args, torch, dist, rank, world, device):
    """The drop-in boundary on N GPUs: rank 0 alone builds the C2 scene through the ANARI C API on a device whose
    `cudaDevices` parameter lists all N GPUs of this job (multiGpuMode sortLast: one process, peer access, z-slabs made
    by the field's finalize, one fused frame kernel per GPU behind anariRenderFrame, the display GPU's frame behind
    anariMapFrame) and runs the same per-step sequence as the N = 1 e2e — camera parameters, commit, render, wait, map
    the colour channel to the host.  The other ranks hold no scene any more and wait at the barrier."""
    res = None
    # The waiting ranks must not wait ON their GPU: an NCCL barrier is a kernel that spins on the device, next to which
    # the fused frame kernel (a persistent grid that needs every CTA resident) crawls — 343 frames/s at N = 2 with it.
    # A gloo group gives a barrier that blocks on the host.
    torch.cuda.synchronize()
    cpu_group = dist.new_group(backend="gloo")
    if rank == 0:
        vol = e = None
        try:
            vol = make_scene(args, torch, device)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e = AnariE2E(args, torch, device, vol, 0, 1, "single", gpus=list(range(world)))
            K = max(args.steps, 20)
            e.prepare(K + 3)
            for i in range(3):
                e.step(i)
            setup_s = time.perf_counter() - t0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(K):
                e.step(i)
            dt = time.perf_counter() - t0
            ndev = e.d.get_property(e.d.handle, "cudaDeviceCount", e.A.INT32)
            hb = e.bytes_per_step()
            res = {"value": K / dt, "unit": "frames/s", "steps": K, "n_gpus_in_device": int(ndev), "setup_s": setup_s,
                   "h2d_bytes_per_step": hb[0], "d2h_bytes_per_step": hb[1],
                   "what": "ONE process, ANARI C API of libanari_library_visrtx_b200.so with device parameters "
                           f"cudaDevices=0..{world - 1}, multiGpuMode=sortLast: anariSetParameter(camera) + "
                           "anariCommitParameters + anariRenderFrame + anariFrameReady(WAIT) + anariMapFrame("
                           "channel.color -> host) per step, wall clock; parity of this path against one GPU: "
                           "tests/test_gpu_anari_multigpu.py"}
            e.close()
            e = None
        except Exception as ex:  # a secondary measurement must never take the headline line down
            res = f"unavailable: {type(ex).__name__}: {ex}"
        finally:
            del vol, e
            # the multi-GPU device switches the CUDA current device while it works: hand this process' own GPU back to
            # everything that follows (torch and the C-ABI both act on the current device)
            try:
                C.CDLL("libcudart.so").cudaSetDevice(C.c_int(device.index))
            except OSError:
                pass
            torch.cuda.set_device(device)
            torch.cuda.empty_cache()
    dist.barrier(group=cpu_group)
    dist.destroy_process_group(cpu_group)
    return res


def measure_c3_sort_first(base_args, torch, dist, capi, rank, world, device, stream):
    """BASELINE config C3 — 2048^3 UFIXED16 sparse shells at 3840x2160, macrocell skipping — rendered sort-first on all
    ranks of this job (field replicated, tile rows interleaved, colour stored straight into the display rank's frame
    through a peer pointer), next to the same frame rendered by the display rank alone from its own replica: the
    assembled frame must be bit-identical, and the ratio of the two rates is the sort-first scaling at this N."""
    import argparse
    from visrtx_b200 import multigpu
    a = argparse.Namespace(**vars(base_args))
    a.config, a.skip, a.field, a.rate, a.mode = "c3", -1, "ml", 0.5, "sort-first"
    a = apply_preset(a)
    n, W, H = a.size, a.width, a.height
    npx = W * H
    FMT, INTEG, BG = capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, (0.1, 0.1, 0.1, 1.0)
    t0 = time.perf_counter()
    vol = make_scene(a, torch, device)
    field = create_field(a, capi, vol, stream)
    torch.cuda.synchronize()
    del vol
    torch.cuda.empty_cache()
    tf = capi.tf_discretize(color=scene_colormap(a))
    volume = capi.Volume.create(field, tf, (0.0, 1.0), a.unit_distance, 0, stream)
    inst, ninst = capi.make_instances([volume], None, [0])
    cam, _ = orbit(a)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    K = 30
    res = {"workload": workload_name(a), "unit": "frames/s", "n_gpus": world, "steps": K, "macrocell_skipping": bool(a.skip)}
    bands = {}
    frame0 = None
    for band in (1, 8):  # tile rows handed out singly / in bands of 8 rows (32 pixel rows)
        drv = multigpu.SortFirst(capi, torch, dist, rank, world, device, W, H, inst, ninst, FMT, INTEG, a.rate, BG,
                                 skip=bool(a.skip), tile_band=band, host_mirror=False)
        drv.render(0, cam, stream)
        torch.cuda.synchronize()
        dist.barrier()
        f0 = drv.color_tensor() if rank == 0 else None
        torch.cuda.synchronize()
        dist.barrier()
        for i in range(5):
            drv.render(1 + i, cam, stream)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            drv.render(6 + i, cam, stream)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bands[band] = float(t.item())
        if band == 1:
            frame0 = f0
        drv.close()
    best = min(bands, key=bands.get)
    # the display rank alone, whole frame from its own replica
    single_ms, identical = None, None
    if rank == 0:
        accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
        color = torch.zeros(npx, dtype=torch.int32, device=device)
        depth = torch.zeros(npx, dtype=torch.float32, device=device)
        fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
        mk = lambda fid: capi.frame_params(W, H, FMT, INTEG, fid, -1, 1, a.rate, BG, skip=bool(a.skip))
        capi.render(mk(0), cam, inst, ninst, fb, stream)
        torch.cuda.synchronize()
        identical = bool(torch.equal(color, frame0))
        for i in range(5):
            capi.render(mk(1 + i), cam, inst, ninst, fb, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            capi.render(mk(6 + i), cam, inst, ninst, fb, stream)
        e1.record()
        torch.cuda.synchronize()
        single_ms = e0.elapsed_time(e1) / K
    dist.barrier()
    volume.destroy()
    field.destroy()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    res.update({"value": 1000.0 / bands[best], "ms_per_step": bands[best], "tile_band": best,
                "ms_per_step_by_tile_band": {str(k): v for k, v in bands.items()},
                "single_gpu_value": 1000.0 / single_ms, "single_gpu_ms": single_ms,
                "speedup_over_one_gpu": single_ms / bands[best], "setup_s": setup_s,
                "parity_vs_single": {"bit_identical": identical, "pass": identical,
                                     "what": f"frame 0 assembled by {world} GPUs (sort-first, peer stores into the display "
                                             "rank's frame) == frame 0 rendered by the display rank alone, every sRGB8 "
                                             "pixel; the same scene against O-gpu: extra.configs.c3.parity of the N = 1 line"},
                "what": "field replicated on every GPU, tile rows interleaved across ranks (the best of tileBand 1 and 8 "
                        "is the value), one frame kernel per GPU storing colour into the display rank's frame over "
                        "NVLink, a device-side barrier per frame; CUDA events, max over ranks"})
    return res


def parity_vs_single(args, torch, dist, capi, driver, cam, rank, world, device, stream, mode):
    """N > 1: the frame the N GPUs assembled against the SAME frame rendered by one GPU holding the whole volume
    (dvr_render on the display rank's GPU).  Only when the whole volume fits beside the rank's own share."""
    n, W, H = args.size, args.width, args.height
    npx = W * H
    need = n ** 3 * voxel_bytes(args) * 2.2
    driver.stream_to_host(False)
    driver.render(0, cam, stream)  # frame 0 (accumulation reset) by all ranks
    torch.cuda.synchronize()
    dist.barrier()
    res = None
    if rank == 0:
        free_b, _ = torch.cuda.mem_get_info()
        if args.field == "fog" or need > free_b:
            res = {"pass": None, "what": f"not run: the whole {n}^3 volume does not fit one GPU next to this rank's share"}
        else:
            multi = driver.color_tensor()
            vol = make_scene(args, torch, device)
            field = create_field(args, capi, vol, stream)
            del vol
            tf = capi.tf_discretize(color=scene_colormap(args))
            v = capi.Volume.create(field, tf, (0.0, 1.0), args.unit_distance, 0, stream)
            inst, ninst = capi.make_instances([v], None, [0])
            accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
            color = torch.zeros(npx, dtype=torch.int32, device=device)
            depth = torch.zeros(npx, dtype=torch.float32, device=device)
            fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
            p = capi.frame_params(W, H, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, 0, -1, 1, args.rate,
                                  (0.1, 0.1, 0.1, 1.0), skip=bool(args.skip))
            capi.render(p, cam, inst, ninst, fb, stream)
            torch.cuda.synchronize()
            et = None
            if mode == "sort-last":
                # which rays the one-pass march terminates early: the same frame over a background of alpha 0 leaves the
                # volume's own opacity in the accumulation buffer's alpha
                accum2 = torch.zeros((npx, 4), dtype=torch.float32, device=device)
                color2 = torch.zeros(npx, dtype=torch.int32, device=device)
                fb2 = capi.frame_buffers(accum2.data_ptr(), color2.data_ptr(), 0)
                p2 = capi.frame_params(W, H, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, 0, -1, 1,
                                       args.rate, (0.1, 0.1, 0.1, 0.0), skip=bool(args.skip))
                capi.render(p2, cam, inst, ninst, fb2, stream)
                torch.cuda.synchronize()
                et = accum2[:, 3] >= 0.99
            res = parity_block(torch, 4, multi, color, what=f"frame 0 assembled by {world} GPUs ({mode}) vs dvr_render of "
                               "the whole volume on one GPU, same camera and Philox streams, sRGB8 colour",
                               early_terminated=et)
            v.destroy()
            field.destroy()
    dist.barrier()
    return res


def measure_c4_scaling(base_args, torch, dist, capi, rank, world, device, stream):
    """BASELINE config C4 — 4096^3 f32 (256 GiB), bricked sort-last — on 2, 4 and 8 of this job's ranks, each time
    from scratch: slabs of equal work generated in HBM in chunks, one fused launch per GPU and frame.  Frame 0 of
    every sub-run is kept on rank 0 and compared with the previous one: the volume fits no single GPU, so the images
    of different partitions are checked against each other (each is the same global sample lattice)."""
    import argparse
    from visrtx_b200 import multigpu
    a = argparse.Namespace(**vars(base_args))
    a.config, a.skip, a.field, a.rate = "c4", -1, "ml", 0.5
    a = apply_preset(a)
    n, W, H = a.size, a.width, a.height
    npx = W * H
    res = {"workload": workload_name(a), "unit": "frames/s", "runs": {}}
    prev_frame, prev_n, prev_et = None, None, None
    groups = {sub: dist.new_group(list(range(sub))) for sub in (2, 4, 8) if sub <= world}
    for sub, group in groups.items():
        dist.barrier()
        entry = None
        if rank < sub:
            t0 = time.perf_counter()
            lo, hi = scene_bounds(a)
            cam, pose0 = orbit(a)
            ranges0 = multigpu.view_balanced_slab_ranges(n, sub, lo, hi, pose0.position)
            margin = multigpu.slab_margin(n, sub) if base_args.balance else 0
            limits = multigpu.creation_ranges(ranges0, n, margin)
            z0, z1 = limits[rank]
            r0, r1 = multigpu.resident_range(z0, z1, n)
            field = capi.Field.create_slab(0, True, scene_dtype(a), (n, n, n), z0, z1, (0, 0, 0), (1, 1, 1),
                                           capi.DVR_FILTER_LINEAR, stream)
            field.set_owned_slices(*ranges0[rank])
            for zc in range(r0, r1, 32):
                ze = min(zc + 32, r1)
                part = make_scene(a, torch, device, z_begin=zc, z_end=ze)
                field.upload_slices(part.data_ptr(), True, zc - r0, ze - zc, stream)
                torch.cuda.synchronize()
                del part
            field.build_macrocells(stream)
            tf = capi.tf_discretize(color=scene_colormap(a))
            volume = capi.Volume.create(field, tf, (0.0, 1.0), a.unit_distance, 0, stream)
            inst, _ = capi.make_instances([volume], None, [0])
            torch.cuda.synchronize()
            setup_s = time.perf_counter() - t0
            drv = multigpu.SortLast(capi, torch, dist, rank, sub, device, W, H, inst, 0, 0, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB,
                                    capi.DVR_INTEGRATOR_DEFAULT, a.rate, (0.1, 0.1, 0.1, 1.0), skip=False, group=group)
            balance = None
            if base_args.balance:
                final, _, hist = drv.calibrate(field, ranges0, limits, cam, stream, rounds=2, frames=4)
                balance = {"margin_slices": margin, "initial": [list(r) for r in ranges0], "final": [list(r) for r in final],
                           "rounds": hist}
            drv.render(0, cam, stream)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            frame0 = drv.color_tensor() if rank == 0 else None
            torch.cuda.synchronize()   # the copy is done before any rank's next frame can store into the display frame
            dist.barrier(group=group)
            # the same frame over a background of alpha 0: the alpha byte is then the volume's own opacity, which tells
            # the early-terminated rays (>= 0.99) apart for the parity block
            bg_keep = drv.background
            drv.background = (bg_keep[0], bg_keep[1], bg_keep[2], 0.0)
            drv._hot["p"] = drv.params(0)
            drv.render(0, cam, stream)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            et0 = ((drv.color_tensor().view(torch.uint8).view(-1, 4)[:, 3] >= 252) if rank == 0 else None)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            drv.background = bg_keep
            drv._hot["p"] = drv.params(0)
            K = 30
            for i in range(5):
                drv.render(1 + i, cam, stream)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                drv.render(6 + i, cam, stream)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / K, setup_s], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            err = drv.check_errors()
            if rank == 0:
                entry = {"value": 1000.0 / float(t[0].item()), "ms_per_step": float(t[0].item()), "steps": K,
                         "setup_s": float(t[1].item()), "slab_gib_per_gpu": (r1 - r0) * n * n * 4 / 2 ** 30,
                         "spin_timeouts": bool(err), "slab_balance": balance}
                if prev_frame is not None:
                    entry[f"parity_vs_n{prev_n}"] = parity_block(
                        torch, 4, frame0, prev_frame, what=f"frame 0 on {sub} GPUs vs frame 0 on {prev_n} GPUs (the volume "
                        "fits no single GPU: partitions are checked against each other; early-terminated rays = "
                        "composited opacity >= 252/255 in either partition)", early_terminated=et0 | prev_et)
                prev_frame, prev_n, prev_et = frame0, sub, et0
            drv.close()
            volume.destroy()
            field.destroy()
            torch.cuda.empty_cache()
        dist.barrier()
        if rank == 0:
            res["runs"][f"n{sub}"] = entry
    if rank == 0 and "n2" in res["runs"] and "n8" in res["runs"]:
        res["speedup_8_over_2"] = res["runs"]["n8"]["value"] / res["runs"]["n2"]["value"]
    return res


def measure_secondary_config(base_args, name, torch, device, stream):
    """Another BASELINE config at N = 1, after the headline: device-timed frames/s, roofline from an instrumented
    launch, e2e through the ANARI C API, and parity against O-gpu on frame 0 — driver-observed versions of the numbers
    DESIGN.md quotes for C3 / C5."""
    import argparse
    from visrtx_b200 import capi
    a = argparse.Namespace(**vars(base_args))
    a.config, a.skip, a.field, a.rate, a.nvdb_codec = name, -1, "ml", 0.5, "float"
    a = apply_preset(a)
    W, H, npx = a.width, a.height, a.width * a.height
    t0 = time.perf_counter()
    vol = make_scene(a, torch, device)
    field = create_field(a, capi, vol, stream)
    torch.cuda.synchronize()
    tf = capi.tf_discretize(color=scene_colormap(a))
    v = capi.Volume.create(field, tf, (0.0, 1.0), a.unit_distance, 0, stream)
    inst, ninst = capi.make_instances([v], None, [0])
    cam, _ = orbit(a)
    setup_s = time.perf_counter() - t0
    accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
    color = torch.zeros(npx, dtype=torch.int32, device=device)
    depth = torch.zeros(npx, dtype=torch.float32, device=device)
    fb = capi.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
    mk = lambda fid: capi.frame_params(W, H, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, fid, -1, 1,
                                       a.rate, (0.1, 0.1, 0.1, 1.0), skip=bool(a.skip))
    st = torch.zeros(4, dtype=torch.int64, device=device)
    capi.render_instrumented(mk(0), cam, inst, ninst, fb, st.data_ptr(), stream)
    torch.cuda.synchronize()
    samples, skipped, rays_hit, cells = st.tolist()
    for i in range(5):
        capi.render(mk(i), cam, inst, ninst, fb, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 64 if name == "c5" else 30  # C5: the 64-frame progressive accumulation of BASELINE.json
    e0.record()
    for i in range(K):
        capi.render(mk(i), cam, inst, ninst, fb, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    peak, peak_src = measured_peak()
    note = ("sparse / early-terminating workloads are latency- and issue-bound, not HBM-bound: the fraction is "
            "reported, not claimed as the bound")
    if a.field == "fog":
        # apron bricks: 9^3 floats per non-constant 8^3 cell.  A sample reads at most 8 sectors, but HBM never has to
        # deliver more than the resident field once per frame: the algorithmic bytes are the smaller of the two.
        resident = int(field.device_bytes())
        sector_cap = samples * 8 * 32
        if resident < sector_cap:
            bframe, bmodel = resident + npx * 44 + 4096, ("resident field bytes (bricks + table + grid) read once per "
                                                          "frame; the sector cap samples * 8 taps * 32 B is larger")
            note = (f"the resident field ({resident / 2 ** 20:.0f} MiB) is of the order of the 126 MB L2 and is re-read "
                    "every frame, so most fetches are L2 hits: this kernel is bound by L1/L2 latency and issue, not by "
                    "HBM (profiles/r02_c5_brick_march_ncu.md); the HBM fraction is reported, not claimed as the bound")
        else:
            bframe, bmodel = sector_cap + npx * 44 + 4096, "sector cap: samples * 8 taps * 32 B (brick rows are not sector aligned)"
    else:
        bframe, bmodel = bytes_per_frame(a, cells, samples)
    res = {"workload": workload_name(a), "value": 1000.0 / ms, "unit": "frames/s", "ms_per_step": ms, "steps": K,
           "samples_per_frame": int(samples), "samples_skipped": int(skipped), "gsamples_per_s": samples / ms / 1e6,
           "macrocell_skipping": bool(a.skip), "setup_s": setup_s,
           "roofline": {"bound": "hbm", "achieved": bframe / ms / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": bframe / ms / 1e6 / peak, "algorithmic_bytes": bframe, "bytes_model": bmodel,
                        "macrocells_touched": int(cells), "kernel_ms": ms, "peak_source": peak_src,
                        "note": note}}
    try:
        (ref_color, ref_depth), ref_fps = ref_gpu_frame0(a, torch, vol, 10)
        capi.render(mk(0), cam, inst, ninst, fb, stream)
        torch.cuda.synchronize()
        res["parity"] = parity_block(torch, 4, color, ref_color, depth, ref_depth,
                                     what="frame 0: dvr_render vs O-gpu (reference device code), same scene and camera")
        res["ref_gpu_fps"] = ref_fps
        del ref_color, ref_depth
    except Exception as e:
        res["parity"] = {"pass": None, "what": f"unavailable: {e}"}
    v.destroy()
    field.destroy()
    del vol, accum, color, depth
    torch.cuda.empty_cache()
    try:
        import types
        vol2 = make_scene(a, torch, device)
        e2e = AnariE2E(a, torch, device, vol2, 0, 1, "single")
        e2e.prepare(K + 3)
        for i in range(3):
            e2e.step(i)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for i in range(K):
            e2e.step(i)
        torch.cuda.synchronize()
        res["e2e"] = {"value": K / (time.perf_counter() - t1), "unit": "frames/s",
                      "h2d_bytes_per_step": e2e.bytes_per_step()[0], "d2h_bytes_per_step": e2e.bytes_per_step()[1]}
        e2e.close()
        del vol2
        torch.cuda.empty_cache()
    except Exception as e:
        res["e2e"] = {"value": None, "what": f"unavailable: {e}"}
    return res


def measure_variants(args, torch, capi, scenes, field, cam, inst, ninst, fb, stream, stats_t):
    """Secondary numbers (same scene): other sampling rates and the opaque BASELINE.md-literal TF."""
    res = {}
    tf = capi.tf_discretize(color=scene_colormap(args))
    for name, rate, ud in (("rate0.125_ud256", 0.125, args.unit_distance), ("rate1.0_ud256", 1.0, args.unit_distance),
                           ("rate0.5_ud1_opaque", 0.5, 1.0)):
        v = capi.Volume.create(field, tf, (0.0, 1.0), ud, 0, stream)
        ins, nn = capi.make_instances([v], None, [0])
        mk = lambda fid: capi.frame_params(args.width, args.height, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB,
                                           capi.DVR_INTEGRATOR_DEFAULT, fid, -1, 1, rate, (0.1, 0.1, 0.1, 1.0))
        capi.render_instrumented(mk(0), cam, ins, nn, fb, stats_t.data_ptr(), stream)
        torch.cuda.synchronize()
        samples, _, _, cells = stats_t.tolist()
        for i in range(3):
            capi.render(mk(i), cam, ins, nn, fb, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            capi.render(mk(3 + i), cam, ins, nn, fb, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        bf, _ = bytes_per_frame(args, cells, samples)
        res[name] = {"fps": 1000.0 / ms, "ms": ms, "samples_per_frame": samples, "gsamples_per_s": samples / ms / 1e6,
                     "macrocells_touched": cells, "algorithmic_GBps": bf / ms / 1e6}
        v.destroy()
    return res


DPT_OPACITY = (0.0, 0.02)  # TF alpha ramp of the dpt variant: mean free path >= 25 voxels, multiple scattering


def measure_time_varying(args, torch, capi, field, cam, fb, stream, vol_dev):
    """A field that changes every frame (in-situ): per step the whole f32 volume is re-finalised from
    device memory (dvr_field_update_structured: upload + macrocell ranges in one pass), the volume re-derives its
    majorants, and one frame is marched.  Wall clock, result left on the device."""
    n = args.size
    tf = capi.tf_discretize(color=scene_colormap(args))
    v = capi.Volume.create(field, tf, (0.0, 1.0), args.unit_distance, 0, stream)
    ins, nn = capi.make_instances([v], None, [0])
    p = capi.frame_params(args.width, args.height, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB, capi.DVR_INTEGRATOR_DEFAULT, 0,
                          -1, 1, args.rate, (0.1, 0.1, 0.1, 1.0))

    def step():
        field.update_structured(vol_dev.data_ptr(), True, capi.DVR_FLOAT32, (0, 0, 0), (1, 1, 1), stream)
        v.update(tf, (0.0, 1.0), args.unit_distance, 0, stream)
        capi.render(p, cam, ins, nn, fb, stream)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 20
    v.destroy()
    return {"fps": 1000.0 / ms, "ms_per_update_and_frame": ms, "bytes_refreshed_per_step": n ** 3 * 4,
            "what": "field refresh (4.3 GB f32 from device memory) + volume majorants + one frame, wall clock"}


def measure_dpt(args, torch, capi, field, cam, fb, stream, vol_dev):
    """The `dpt` renderer (delta tracking, maxDepth 5) on the same field/camera, next to O-gpu running the
    reference's tracker over the same majorant grid."""
    import numpy as np
    tf = capi.tf_discretize(color=scene_colormap(args), opacity=np.asarray(DPT_OPACITY, np.float32))
    v = capi.Volume.create(field, tf, (0.0, 1.0), args.unit_distance, 0, stream)
    ins, nn = capi.make_instances([v], None, [0])
    mk = lambda fid: capi.frame_params(args.width, args.height, capi.DVR_FORMAT_UFIXED8_RGBA_SRGB,
                                       capi.DVR_INTEGRATOR_DPT, fid, -1, 1, args.rate, (0.1, 0.1, 0.1, 1.0))
    dims, maj_ptr = v.dda_majorants(stream)
    for i in range(3):
        capi.render(mk(i), cam, ins, nn, fb, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        capi.render(mk(3 + i), cam, ins, nn, fb, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    res = {"fps": 1000.0 / ms, "ms": ms, "max_depth": 5, "tf_opacity_ramp": list(DPT_OPACITY),
           "grid_dims": list(dims)}
    try:
        lib, sc, (rf, rv) = _refgpu_objects(args, torch, vol_dev, tf=tf, grid=(dims, maj_ptr))
        for i in range(2):
            lib.refgpu_render(C.byref(mk(i)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
        torch.cuda.synchronize()
        e0.record()
        for i in range(10):
            lib.refgpu_render(C.byref(mk(2 + i)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
        e1.record()
        torch.cuda.synchronize()
        res["ref_gpu_fps"] = 10 * 1000.0 / e0.elapsed_time(e1)
        lib.refgpu_scene_destroy(sc)
        lib.refgpu_volume_destroy(rv)
        lib.refgpu_field_destroy(rf)
    except Exception as e:
        res["ref_gpu_fps"] = f"unavailable: {e}"
    v.destroy()
    return res


# ---------------------------------------------------------------------------------------------------------
def _refgpu_objects(args, torch, vol_dev, tf=None, grid=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from visrtx_b200 import pods
    lib = ob.refgpu()
    n = args.size
    f = C.c_void_p()
    if args.field == "fog":
        rc = lib.refgpu_field_create_nvdb(vol_dev.ctypes.data_as(C.c_void_p), C.c_size_t(vol_dev.nbytes), C.byref(f))
    else:
        rc = lib.refgpu_field_create(C.c_void_p(vol_dev.data_ptr()), C.c_int(scene_dtype(args)),
                                     (C.c_uint32 * 3)(n, n, n), (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(1, 1, 1),
                                     C.c_int(0), C.byref(f))
    assert rc == 0, lib.refgpu_last_error()
    if tf is None:
        tf = oracle_tf(args)
    v = C.c_void_p()
    rc = lib.refgpu_volume_create(f, tf.ctypes.data_as(C.c_void_p), (C.c_float * 2)(0, 1), C.c_float(args.unit_distance),
                                  C.c_uint32(0), C.byref(v))
    assert rc == 0, lib.refgpu_last_error()
    if grid is not None:  # delta-tracking grid (dims, device or host pointer to the majorants)
        rc = lib.refgpu_volume_set_grid(v, (C.c_int * 3)(*grid[0]), C.c_void_p(grid[1]))
        assert rc == 0, lib.refgpu_last_error()
    inst = (ob.RefInstance * 1)()
    inst[0].volume = v
    inst[0].worldToObject = (C.c_float * 12)(*pods.IDENTITY_3X4)
    inst[0].instanceId = 0
    sc = C.c_void_p()
    rc = lib.refgpu_scene_create(inst, C.c_int(1), C.byref(sc))
    assert rc == 0, lib.refgpu_last_error()
    return lib, sc, (f, v)


def ref_gpu_frame0(args, torch, vol_dev, steps=0):
    """O-gpu (the reference's device code) on the same scene: frame 0 of the benchmark camera (colour + depth, kept on
    the device for the parity block) and, with steps > 0, its device-timed frames/s."""
    from visrtx_b200 import pods
    lib, sc, (rf, rv) = _refgpu_objects(args, torch, vol_dev)
    npx = args.width * args.height
    dev = torch.device("cuda", torch.cuda.current_device())
    accum = torch.zeros((npx, 4), dtype=torch.float32, device=dev)
    color = torch.zeros(npx, dtype=torch.int32, device=dev)
    depth = torch.zeros(npx, dtype=torch.float32, device=dev)
    fb = pods.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
    cam, _ = orbit(args, oracle_side=True)
    stream = torch.cuda.current_stream().cuda_stream
    mk = lambda fid: pods.frame_params(args.width, args.height, pods.DVR_FORMAT_UFIXED8_RGBA_SRGB,
                                       pods.DVR_INTEGRATOR_DEFAULT, fid, -1, 1, args.rate, (0.1, 0.1, 0.1, 1.0))
    rc = lib.refgpu_render(C.byref(mk(0)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
    assert rc == 0, lib.refgpu_last_error()
    torch.cuda.synchronize()
    frame0 = (color.clone(), depth.clone())
    fps = None
    if steps > 0:
        for i in range(3):
            lib.refgpu_render(C.byref(mk(1 + i)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            lib.refgpu_render(C.byref(mk(4 + i)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
        e1.record()
        torch.cuda.synchronize()
        fps = steps * 1000.0 / e0.elapsed_time(e1)
    lib.refgpu_scene_destroy(sc)
    lib.refgpu_volume_destroy(rv)
    lib.refgpu_field_destroy(rf)
    return frame0, fps


REFERENCE_KIND = ("O-gpu: VisRTX device headers (gpu/volumeIntegration.h, sampleSpatialField.h, gpu_util.h ...) compiled "
                  "for sm_100a from /root/reference, one thread per pixel, OptiX volume-BVH trace replaced by the slab "
                  "test; VisRTX + OptiX proper cannot be built here (no OptiX headers, no ANARI-SDK)")


def bench_config(args, mode, world):
    """`config` of the JSON line: the SAME keys and values in both arms (the driver compares them)."""
    n = args.size
    return {
        "workload": workload_name(args),
        "parallelism": mode + (f"x{world}" if world > 1 else ""),
        "l2": ("NanoVDB grid > 126 MB L2; no flush" if args.field == "fog" else
               f"input volume {n ** 3 * voxel_bytes(args) / 2 ** 30:.1f} GiB >> 126 MB L2; no flush needed"),
        "macrocell_skipping": bool(args.skip),
    }


def metric_name(args):
    return ("DVR frames/s, 1080p, 1024^3 f32 volume (Gsamples/s in extra)" if args.config == "c2"
            else f"DVR frames/s, config {args.config}")


def run_reference(args, torch, dist, rank, world):
    """Reference arm: O-gpu (reference device headers) on ONE GPU; rank 0 only.  Nothing of the product is loaded:
    parameter blocks come from visrtx_b200.pods (plain ctypes structs), camera and transfer function from the oracle
    libraries, the kernels from oracle/_ref/libref_gpu_dvr.so."""
    from visrtx_b200 import pods
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(device)
    n, W, H = args.size, args.width, args.height
    npx = W * H
    vol = make_scene(args, torch, device)
    mode = "single" if world == 1 else ("sort-last" if args.mode == "auto" else args.mode)
    base = {"impl": "reference", "metric": metric_name(args),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (analytic field generated in HBM: " + args.field + ")",
            "config": bench_config(args, mode, world)}
    if not ob.have_ref_gpu():
        # O-gpu did not travel: time the CPU port on a band of rows and scale by rows (the centre band is the most
        # expensive part of the frame, so this under-states the CPU)
        rows = args.cpu_rows or 8
        cb = cpu_baseline(args, torch, vol, None, rows=rows)
        fps = cb["value"]
        base.update({"value": fps, "ms_per_step": 1000.0 / fps, "cpu_baseline": cb, "gpu_launches": 0,
                     "reference_kind": "O-cpu port (oracle/liboracle_dvr.so): oracle/_ref/libref_gpu_dvr.so is absent",
                     "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return base

    lib, sc, _ = _refgpu_objects(args, torch, vol)
    accum = torch.zeros((npx, 4), dtype=torch.float32, device=device)
    color = torch.zeros(npx, dtype=torch.int32, device=device)
    depth = torch.zeros(npx, dtype=torch.float32, device=device)
    fb = pods.frame_buffers(accum.data_ptr(), color.data_ptr(), depth.data_ptr())
    host_color = torch.empty(npx, dtype=torch.int32, pin_memory=True)
    stream = torch.cuda.current_stream().cuda_stream
    mk = lambda fid: pods.frame_params(W, H, pods.DVR_FORMAT_UFIXED8_RGBA_SRGB, pods.DVR_INTEGRATOR_DEFAULT, fid, -1, 1,
                                       args.rate, (0.1, 0.1, 0.1, 1.0))
    cam, _ = orbit(args, oracle_side=True)
    # clocks are sampled from the warm-up to the end of the e2e loop (the timed region alone can be shorter than
    # one 200 ms nvidia-smi period)
    sampler = ClockSampler(device.index)
    sampler.start()
    for i in range(args.warmup):
        lib.refgpu_render(C.byref(mk(i)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        lib.refgpu_render(C.byref(mk(args.warmup + i)), C.byref(cam), sc, C.byref(fb), C.c_void_p(stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps

    e2e_cams = [orbit(args, az_deg=30.0 + 0.05 * i, oracle_side=True)[0] for i in range(args.steps + 3)]
    p0 = mk(0)

    def e2e_step(i):
        cam_i = e2e_cams[i % len(e2e_cams)]
        lib.refgpu_render(C.byref(p0), C.byref(cam_i), sc, C.byref(fb), C.c_void_p(stream))
        host_color.copy_(color, non_blocking=True)
        torch.cuda.synchronize()

    for i in range(3):
        e2e_step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    e2e_fps = args.steps / (time.perf_counter() - t0)
    clocks = sampler.stop()
    base.update({
        "value": 1000.0 / ms, "ms_per_step": ms,
        "reference_kind": REFERENCE_KIND,
        "gpu_launches": args.steps * 4,
        "cpu_baseline": {"value": 1000.0 / ms, "unit": "frames/s", "cores": 0, "kind": "reference",
                         "sample": "full frames on the B200 (the reference is GPU code; see reference_kind)"},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 4 * 1024, "d2h_bytes_per_step": npx * 4,
                "what": "moved camera (constant upload) + refgpu_render + BLOCKING device->host copy of the colour buffer "
                        "to pinned memory every step, wall clock (the product arm instead streams the colour into a "
                        "pinned mirror during the launch)"},
        "clocks": clocks,
    })
    return base


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the DVR path has no CPU fallback"}))
        sys.exit(1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        # the reference has no multi-GPU path: rank 0 alone runs it, the other ranks leave without work
        if rank == 0:
            print(json.dumps(run_reference(args, torch, None, 0, world)), flush=True)
        return
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    try:
        out = run_ours(args, torch, dist, rank, world)
        if rank == 0:
            print(json.dumps(out), flush=True)
    finally:
        if dist is not None and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
