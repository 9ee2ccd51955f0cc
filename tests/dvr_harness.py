"""Shared scene harness for the parity tests: one scene description, three renderers.

  render_cuda()   the product: libdvr_b200.so through the C-ABI (visrtx_b200.capi), buffers in HBM (torch)
  render_oracle() O-cpu  (oracle/liboracle_dvr.so)
  render_refgpu() O-gpu  (oracle/_ref/libref_gpu_dvr.so, the reference's own device headers)

Every renderer returns a dict of numpy arrays: color, accum, depth, primId, objId, instId (+albedo,
normal when requested).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

import oracle_binding as ob
from visrtx_b200 import capi, scenes

_NP_TYPES = {
    capi.DVR_FLOAT32: np.float32, capi.DVR_UFIXED8: np.uint8, capi.DVR_FIXED8: np.int8,
    capi.DVR_UFIXED16: np.uint16, capi.DVR_FIXED16: np.int16, capi.DVR_FLOAT64: np.float64,
    capi.DVR_FLOAT16: np.float16,
}


def normalized_float(vox: np.ndarray, data_type: int) -> np.ndarray:
    """What cudaReadModeNormalizedFloat hands to the filter (CUDA programming guide, texture read modes)."""
    if data_type == capi.DVR_UFIXED8:
        return (vox.astype(np.float32) / np.float32(255.0)).astype(np.float32)
    if data_type == capi.DVR_FIXED8:
        return np.maximum(vox.astype(np.float32) / np.float32(127.0), np.float32(-1.0)).astype(np.float32)
    if data_type == capi.DVR_UFIXED16:
        return (vox.astype(np.float32) / np.float32(65535.0)).astype(np.float32)
    if data_type == capi.DVR_FIXED16:
        return np.maximum(vox.astype(np.float32) / np.float32(32767.0), np.float32(-1.0)).astype(np.float32)
    return vox.astype(np.float32)


@dataclass
class VolumeDesc:
    voxels: np.ndarray  # (z,y,x)
    data_type: int = capi.DVR_FLOAT32
    origin: tuple = (0.0, 0.0, 0.0)
    spacing: tuple = (1.0, 1.0, 1.0)
    nearest: bool = False
    tf: Optional[np.ndarray] = None  # (256,4)
    value_range: tuple = (0.0, 1.0)
    unit_distance: float = 1.0
    vol_id: int = 0xFFFFFFFF
    inst_id: int = 0xFFFFFFFF
    world_to_object: Optional[tuple] = None
    nvdb: Optional[np.ndarray] = None  # serialized NanoVDB float grid (uint8): the field is a "nanovdb" field

    @property
    def dims(self):
        nz, ny, nx = self.voxels.shape
        return (nx, ny, nz)

    def bounds(self):
        if self.nvdb is not None:
            wb = self.nvdb[560:608].view(np.float64)
            return wb[:3].astype(np.float32), wb[3:].astype(np.float32)
        lo = np.asarray(self.origin, dtype=np.float32)
        hi = lo + (np.asarray(self.dims, dtype=np.float32) - np.float32(1.0)) * np.asarray(self.spacing, np.float32)
        return lo, hi


@dataclass
class SceneDesc:
    volumes: List[VolumeDesc]
    width: int
    height: int
    camera: capi.DvrCamera
    fmt: int = capi.DVR_FORMAT_UFIXED8_RGBA_SRGB
    integrator: int = capi.DVR_INTEGRATOR_RAYCAST
    volume_sampling_rate: float = 0.125
    background: tuple = (0.1, 0.1, 0.1, 1.0)
    num_iterations: int = 1
    channels: tuple = ("depth", "primId", "objId", "instId")
    max_depth: int = 5  # dpt integrator only
    ambient_radiance: float = 1.0
    occlusion_distance: float = 1e20
    dpt_reference_grid: bool = False  # walk the grid content the reference builds (Q7/Q8)
    # renderer "background" given as an Array2D image: (pixels [h, w, channels], capi.DVR_IMAGE_* component type)
    background_image: Optional[tuple] = None
    # mixed scenes: surfaces / lights as dicts with the ANARI parameter names (pods.surface_descs,
    # pods.scene_params) and the renderer members the surface branch reads
    surfaces: Optional[list] = None
    lights: Optional[list] = None
    ambient_color: tuple = (1.0, 1.0, 1.0)
    surface_ambient_radiance: float = 0.0
    ambient_samples: int = 1
    cull_triangle_backfaces: bool = False

    def scene_params(self, surfaces_handle=None):
        return capi.scene_params(surfaces_handle, self.lights or [], self.ambient_color, self.surface_ambient_radiance,
                                 self.occlusion_distance, self.ambient_samples, self.cull_triangle_backfaces)


def default_scene(n=64, width=256, height=256, rate=0.5, field="ml", **kw) -> SceneDesc:
    """BASELINE config C1 (scaled): Marschner-Lobb on [-1,1]^3, TSD default map, orbit camera."""
    vox = scenes.marschner_lobb_np(n) if field == "ml" else scenes.blobs_np(n)
    sp = 2.0 / (n - 1)
    tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256))
    v = VolumeDesc(vox, origin=(-1.0, -1.0, -1.0), spacing=(sp, sp, sp), tf=tf, unit_distance=sp, vol_id=7,
                   inst_id=3)
    lo, hi = v.bounds()
    pose = scenes.orbit_camera(lo, hi, width, height)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    return SceneDesc([v], width, height, cam, volume_sampling_rate=rate, **kw)


def _alloc_np(scene: SceneDesc):
    n = scene.width * scene.height
    out = {"accum": np.zeros((n, 4), np.float32)}
    out["color"] = np.zeros((n, 4), np.float32) if scene.fmt == capi.DVR_FORMAT_FLOAT32_VEC4 else np.zeros(n, np.uint32)
    if "depth" in scene.channels: out["depth"] = np.zeros(n, np.float32)
    if "primId" in scene.channels: out["primId"] = np.zeros(n, np.uint32)
    if "objId" in scene.channels: out["objId"] = np.zeros(n, np.uint32)
    if "instId" in scene.channels: out["instId"] = np.zeros(n, np.uint32)
    if "albedo" in scene.channels: out["albedo"] = np.zeros((n, 3), np.float32)
    if "normal" in scene.channels: out["normal"] = np.zeros((n, 3), np.float32)
    return out


def _params(scene, frame_id, cb, **kw):
    return capi.frame_params(scene.width, scene.height, scene.fmt, scene.integrator, frame_id, cb,
                             scene.num_iterations, scene.volume_sampling_rate, scene.background,
                             max_depth=scene.max_depth, ambient_radiance=scene.ambient_radiance,
                             occlusion_distance=scene.occlusion_distance,
                             dpt_reference_grid=scene.dpt_reference_grid, **kw)


# ------------------------------------------------------------------------------------------------------------
def _oracle_volumes(scene: SceneDesc, slab=None):
    vols = (ob.OracleVolume * max(len(scene.volumes), 1))()
    keep = []
    for i, v in enumerate(scene.volumes):
        f32 = np.ascontiguousarray(normalized_float(v.voxels, v.data_type))
        tf = np.ascontiguousarray(v.tf, np.float32)
        keep += [f32, tf]
        o = vols[i]
        if v.nvdb is not None:
            blob = np.ascontiguousarray(v.nvdb)
            keep.append(blob)
            o.nvdbGrid = blob.ctypes.data_as(C.c_void_p)
        o.voxels = f32.ctypes.data_as(C.c_void_p)
        o.dims = (C.c_int32 * 3)(*v.dims)
        o.origin = (C.c_float * 3)(*v.origin)
        o.spacing = (C.c_float * 3)(*v.spacing)
        o.filterNearest = 1 if v.nearest else 0
        o.tf = tf.ctypes.data_as(C.c_void_p)
        o.valueRange = (C.c_float * 2)(*v.value_range)
        o.unitDistance = v.unit_distance
        o.id = v.vol_id
        o.worldToObject = (C.c_float * 12)(*(v.world_to_object or capi.IDENTITY_3X4))
        o.instanceId = v.inst_id
        if slab is not None:
            o.zOwnBegin, o.zOwnEnd = slab
    return vols, keep


def oracle_dda_grids(scene: SceneDesc):
    """[(dims, float32[gz,gy,gx])] of O-cpu's delta-tracking grids, one per volume."""
    vols, keep = _oracle_volumes(scene)
    out = []
    for i in range(len(scene.volumes)):
        dims = (C.c_int32 * 3)()
        assert ob.cpu().oracle_dda_majorants(C.byref(vols[i]), dims, None, C.c_size_t(0)) == 0
        a = np.zeros(dims[0] * dims[1] * dims[2], np.float32)
        assert ob.cpu().oracle_dda_majorants(C.byref(vols[i]), dims, a.ctypes.data_as(C.c_void_p),
                                             C.c_size_t(a.size)) == 0
        out.append((tuple(dims), a.reshape(dims[2], dims[1], dims[0])))
    return out


def render_oracle(scene: SceneDesc, frames=1, checkerboard=False, slab=None, state=None, return_samples=False):
    """O-cpu; `frames` successive renderFrame calls with accumulation (Frame::newFrame bookkeeping)."""
    out = state or _alloc_np(scene)
    vols, keep = _oracle_volumes(scene, slab)
    b = ob.OracleBuffers()
    b.colorAccumulation = out["accum"].ctypes.data_as(C.c_void_p)
    b.outColor = out["color"].ctypes.data_as(C.c_void_p)
    for name, key in (("depth", "depth"), ("primId", "primId"), ("objId", "objId"), ("instId", "instId"),
                      ("albedo", "albedo"), ("normal", "normal")):
        if key in out:
            setattr(b, name, out[key].ctypes.data_as(C.c_void_p))
    total = 0
    if scene.background_image is not None:
        st = staged_background(*scene.background_image)
        assert ob.cpu().oracle_set_background_image(st.ctypes.data_as(C.c_void_p), C.c_int(st.shape[2]),
                                                    C.c_int(st.shape[1]), C.c_int(st.shape[0])) == 0
    try:
        for frame_id, cb in frame_sequence(scene, frames, checkerboard):
            p = _params(scene, frame_id, cb)
            s = C.c_uint64()
            rc = ob.cpu().oracle_render(C.byref(p), C.byref(scene.camera), vols, len(scene.volumes), C.byref(b),
                                        C.byref(s), 0, 0)
            assert rc == 0
            total += s.value
    finally:
        if scene.background_image is not None:
            ob.cpu().oracle_set_background_image(None, C.c_int(0), C.c_int(0), C.c_int(0))
    if return_samples:
        return out, total
    return out


def staged_background(pixels: np.ndarray, component_type: int) -> np.ndarray:
    """The RGBA8 texels the renderer's staging pass makes of a background array (utility/CudaImageTexture.cpp:43-58,
    84-101): truncating float -> uint8, fixed-point rescale, sRGB linearisation of EVERY component, 3 -> 4 channels
    with alpha 255.  numpy restatement for O-cpu; [h, w, 1|2|4] uint8."""
    px = np.ascontiguousarray(pixels)
    if px.ndim == 2:
        px = px[:, :, None]
    if component_type == capi.DVR_IMAGE_FLOAT32:
        st = (px.astype(np.float32) * np.float32(255)).astype(np.uint8)
    elif component_type == capi.DVR_IMAGE_UFIXED16:
        st = ((px.astype(np.float32) / np.float32(65535.0)) * np.float32(255)).astype(np.uint8)
    elif component_type == capi.DVR_IMAGE_UFIXED32:
        st = ((px.astype(np.float32) / np.float32(4294967295.0)) * np.float32(255)).astype(np.uint8)
    elif component_type == capi.DVR_IMAGE_SRGB8:
        v = px.astype(np.float32) / np.float32(255.0)
        lin = np.where(v <= np.float32(0.04045), v * np.float32(0.07739938080495356),
                       np.power((v + np.float32(0.055)) * np.float32(0.9478672985781991), np.float32(2.4)))
        st = (lin.astype(np.float32) * np.float32(255)).astype(np.uint8)
    else:
        st = px.astype(np.uint8)
    if st.shape[2] == 3:
        st = np.concatenate([st, np.full(st.shape[:2] + (1,), 255, np.uint8)], axis=2)
    return np.ascontiguousarray(st)


def frame_sequence(scene, frames, checkerboard):
    """(frameID, checkerboardID) per renderFrame call: Frame::newFrame, frame/Frame.cu:590-660."""
    seq = []
    frame_id, cb = 0, (0 if checkerboard else -1)
    for i in range(frames):
        if i > 0:
            if checkerboard:
                frame_id += 1 if cb == 3 else 0
                cb = (cb + 1) & 3
            else:
                frame_id += scene.num_iterations
        seq.append((frame_id, cb))
    return seq


# ------------------------------------------------------------------------------------------------------------
class CudaScene:
    """Device-side objects of a SceneDesc created through the C-ABI."""

    def __init__(self, scene: SceneDesc, slab=None, device="cuda:0"):
        import torch
        self.torch = torch
        self.scene = scene
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.fields, self.volumes = [], []
        for v in scene.volumes:
            vox = np.ascontiguousarray(v.voxels.astype(_NP_TYPES[v.data_type], copy=False))
            filt = capi.DVR_FILTER_NEAREST if v.nearest else capi.DVR_FILTER_LINEAR
            if v.nvdb is not None:
                blob = np.ascontiguousarray(v.nvdb)
                f = capi.Field.create_nanovdb(blob.ctypes.data, blob.nbytes, False)
            elif slab is None:
                f = capi.Field.create_structured(vox.ctypes.data, False, v.data_type, v.dims, v.origin, v.spacing, filt)
            else:
                zb, ze = slab
                z0 = max(zb - 1, 0)
                z1 = min(ze + 1, v.dims[2])
                sub = np.ascontiguousarray(vox[z0:z1])
                f = capi.Field.create_slab(sub.ctypes.data, False, v.data_type, v.dims, zb, ze, v.origin, v.spacing, filt)
            self.fields.append(f)
            self.volumes.append(capi.Volume.create(f, v.tf, v.value_range, v.unit_distance, v.vol_id))
        self.instances, self.n = capi.make_instances(
            self.volumes, [v.world_to_object for v in scene.volumes], [v.inst_id for v in scene.volumes])
        self.bg_image = capi.Image.create(*scene.background_image) if scene.background_image is not None else None
        self.surfaces = capi.Surfaces.create(scene.surfaces) if scene.surfaces else None
        n = scene.width * scene.height
        t = torch
        self.buf = {"accum": t.zeros((n, 4), dtype=t.float32, device=self.device)}
        if scene.fmt == capi.DVR_FORMAT_FLOAT32_VEC4:
            self.buf["color"] = t.zeros((n, 4), dtype=t.float32, device=self.device)
        else:
            self.buf["color"] = t.zeros(n, dtype=t.int32, device=self.device)
        for key, shape, dt in (("depth", (n,), t.float32), ("primId", (n,), t.int32), ("objId", (n,), t.int32),
                               ("instId", (n,), t.int32), ("albedo", (n, 3), t.float32), ("normal", (n, 3), t.float32)):
            if key in scene.channels:
                # poison so that "initialised by the launch" is really tested
                self.buf[key] = t.full(shape, -12345, dtype=dt, device=self.device)
        self.buf["accum"].fill_(777.0)
        g = lambda k: self.buf[k].data_ptr() if k in self.buf else 0
        self.fb = capi.frame_buffers(g("accum"), g("color"), g("depth"), g("primId"), g("objId"), g("instId"),
                                     g("albedo"), g("normal"))

    def render(self, frame_id=0, cb=-1, skip=False, tile_rank=0, tile_ranks=1, stats=False):
        p = _params(self.scene, frame_id, cb, skip=skip, tile_rank=tile_rank, tile_ranks=tile_ranks,
                    background_image=self.bg_image)
        if stats:
            st = self.torch.zeros(4, dtype=self.torch.int64, device=self.device)
            capi.render_instrumented(p, self.scene.camera, self.instances, self.n, self.fb, st.data_ptr())
            self.torch.cuda.synchronize()
            return dict(zip(("samplesTaken", "samplesSkipped", "raysHit", "macrocellsTouched"), st.tolist()))
        if self.surfaces is not None:
            sp, keep = self.scene.scene_params(self.surfaces.handle)
            capi.render_scene(p, self.scene.camera, self.instances, self.n, sp, self.fb)
            del keep
            return None
        capi.render(p, self.scene.camera, self.instances, self.n, self.fb)
        return None

    def dda_grids(self, reference_build=False):
        """[(dims, float32[nz,ny,nx] majorants)] of the delta-tracking grids (built on demand)."""
        out = []
        for v in self.volumes:
            dims, ptr = v.dda_majorants(reference_build=reference_build)
            n = dims[0] * dims[1] * dims[2]
            a = np.zeros(n, np.float32)
            self.torch.cuda.synchronize()
            C.CDLL("libcudart.so").cudaMemcpy(a.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(n * 4), C.c_int(2))
            out.append((dims, a.reshape(dims[2], dims[1], dims[0])))
        return out

    def download(self):
        self.torch.cuda.synchronize()
        out = {}
        for k, v in self.buf.items():
            a = v.cpu().numpy()
            if a.dtype == np.int32:
                a = a.view(np.uint32)
            out[k] = a
        return out

    def destroy(self):
        for v in self.volumes:
            v.destroy()
        for f in self.fields:
            f.destroy()
        if self.bg_image is not None:
            self.bg_image.destroy()
        if self.surfaces is not None:
            self.surfaces.destroy()


def render_cuda(scene: SceneDesc, frames=1, checkerboard=False, skip=False):
    cs = CudaScene(scene)
    try:
        for frame_id, cb in frame_sequence(scene, frames, checkerboard):
            cs.render(frame_id, cb, skip=skip)
        return cs.download()
    finally:
        cs.destroy()


# ------------------------------------------------------------------------------------------------------------
def render_refgpu(scene: SceneDesc, frames=1, checkerboard=False, grids=None, return_grids=False):
    """O-gpu: the reference's device headers (tex3D / tex1D / cuRAND) on the same GPU.
    grids: per volume (dims, majorants) of the delta-tracking grid (dpt integrator only), or the string
    "reference": every volume gets the grid the reference's own UniformGrid code builds for it."""
    import torch
    lib = ob.refgpu()
    fields, vols = [], []
    ref_grids = []
    inst = (ob.RefInstance * max(len(scene.volumes), 1))()
    keep = []
    for i, v in enumerate(scene.volumes):
        vox = np.ascontiguousarray(v.voxels.astype(_NP_TYPES[v.data_type], copy=False))
        keep.append(vox)
        f = C.c_void_p()
        if v.nvdb is not None:
            blob = np.ascontiguousarray(v.nvdb)
            keep.append(blob)
            rc = lib.refgpu_field_create_nvdb(blob.ctypes.data_as(C.c_void_p), C.c_size_t(blob.nbytes), C.byref(f))
        else:
            rc = lib.refgpu_field_create(vox.ctypes.data_as(C.c_void_p), C.c_int(v.data_type),
                                         (C.c_uint32 * 3)(*v.dims), (C.c_float * 3)(*v.origin),
                                         (C.c_float * 3)(*v.spacing), C.c_int(1 if v.nearest else 0), C.byref(f))
        assert rc == 0, lib.refgpu_last_error()
        tf = np.ascontiguousarray(v.tf, np.float32)
        h = C.c_void_p()
        rc = lib.refgpu_volume_create(f, tf.ctypes.data_as(C.c_void_p), (C.c_float * 2)(*v.value_range),
                                      C.c_float(v.unit_distance), C.c_uint32(v.vol_id), C.byref(h))
        assert rc == 0, lib.refgpu_last_error()
        if isinstance(grids, str) and grids == "reference":
            gdims = (C.c_int * 3)()
            rc = lib.refgpu_volume_build_reference_grid(h, gdims, None, C.c_size_t(0))
            assert rc == 0, lib.refgpu_last_error()
            gm = np.zeros(gdims[0] * gdims[1] * gdims[2], np.float32)
            rc = lib.refgpu_volume_build_reference_grid(h, gdims, gm.ctypes.data_as(C.c_void_p), C.c_size_t(gm.size))
            assert rc == 0, lib.refgpu_last_error()
            ref_grids.append((tuple(gdims), gm.reshape(gdims[2], gdims[1], gdims[0])))
        elif grids is not None:
            gd, gm = grids[i]
            gm = np.ascontiguousarray(gm, np.float32)
            rc = lib.refgpu_volume_set_grid(h, (C.c_int * 3)(*gd), gm.ctypes.data_as(C.c_void_p))
            assert rc == 0, lib.refgpu_last_error()
        fields.append(f)
        vols.append(h)
        inst[i].volume = h
        inst[i].worldToObject = (C.c_float * 12)(*(v.world_to_object or capi.IDENTITY_3X4))
        inst[i].instanceId = v.inst_id
    sc = C.c_void_p()
    rc = lib.refgpu_scene_create(inst, C.c_int(len(scene.volumes)), C.byref(sc))
    assert rc == 0, lib.refgpu_last_error()
    ref_img = None
    if scene.background_image is not None:
        px = np.ascontiguousarray(scene.background_image[0])
        if px.ndim == 2:
            px = px[:, :, None]
        ref_img = C.c_void_p()
        rc = lib.refgpu_image_create(px.ctypes.data_as(C.c_void_p), C.c_int(scene.background_image[1]),
                                     C.c_int(px.shape[2]), C.c_uint32(px.shape[1]), C.c_uint32(px.shape[0]),
                                     C.byref(ref_img))
        assert rc == 0, lib.refgpu_last_error()
        lib.refgpu_scene_set_background_image(sc, ref_img)
    if scene.surfaces:
        sarr, skeep = capi.surface_descs(scene.surfaces)
        rc = lib.refgpu_scene_set_surfaces(sc, sarr, C.c_uint32(len(scene.surfaces)))
        assert rc == 0, lib.refgpu_last_error()
        sp, lkeep = scene.scene_params(None)
        rc = lib.refgpu_scene_set_lighting(sc, C.byref(sp))
        assert rc == 0, lib.refgpu_last_error()
        keep += [skeep, lkeep]
    n = scene.width * scene.height
    dev = torch.device("cuda:0")
    buf = {"accum": torch.full((n, 4), 777.0, dtype=torch.float32, device=dev)}
    buf["color"] = (torch.zeros((n, 4), dtype=torch.float32, device=dev)
                    if scene.fmt == capi.DVR_FORMAT_FLOAT32_VEC4 else torch.zeros(n, dtype=torch.int32, device=dev))
    for key, shape, dt in (("depth", (n,), torch.float32), ("primId", (n,), torch.int32), ("objId", (n,), torch.int32),
                           ("instId", (n,), torch.int32), ("albedo", (n, 3), torch.float32),
                           ("normal", (n, 3), torch.float32)):
        if key in scene.channels:
            buf[key] = torch.full(shape, -12345, dtype=dt, device=dev)
    g = lambda k: buf[k].data_ptr() if k in buf else 0
    fb = capi.frame_buffers(g("accum"), g("color"), g("depth"), g("primId"), g("objId"), g("instId"), g("albedo"),
                            g("normal"))
    for frame_id, cb in frame_sequence(scene, frames, checkerboard):
        p = _params(scene, frame_id, cb)
        rc = lib.refgpu_render(C.byref(p), C.byref(scene.camera), sc, C.byref(fb), C.c_void_p(0))
        assert rc == 0, lib.refgpu_last_error()
    torch.cuda.synchronize()
    out = {}
    for k, v in buf.items():
        a = v.cpu().numpy()
        out[k] = a.view(np.uint32) if a.dtype == np.int32 else a
    lib.refgpu_scene_destroy(sc)
    if ref_img is not None:
        lib.refgpu_image_destroy(ref_img)
    for h in vols:
        lib.refgpu_volume_destroy(h)
    for f in fields:
        lib.refgpu_field_destroy(f)
    if return_grids:
        return out, ref_grids
    return out


# ------------------------------------------------------------------------------------------------------------
def unpack_rgba8(u32: np.ndarray) -> np.ndarray:
    return np.stack([(u32 >> s) & 0xFF for s in (0, 8, 16, 24)], axis=-1).astype(np.int32)


def compare_color(a: np.ndarray, b: np.ndarray, fmt: int):
    """max abs per-channel difference (in 1/255 units) and PSNR in dB."""
    if fmt == capi.DVR_FORMAT_FLOAT32_VEC4:
        d = np.abs(a.astype(np.float64) - b.astype(np.float64)) * 255.0
    else:
        d = np.abs(unpack_rgba8(a) - unpack_rgba8(b)).astype(np.float64)
    mse = float(np.mean((d / 255.0) ** 2))
    psnr = 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)
    return float(d.max()), psnr


# ------------------------------------------------------------------------------------------------------------
# The scene zoo shared by the golden-fixture generator (tests/golden/make_golden.py, run on a B200 with
# O-gpu) and the parity tests.  name -> (SceneDesc, frames, checkerboard)
# ------------------------------------------------------------------------------------------------------------
def _rot_y(deg, tx=0.0, ty=0.0, tz=0.0):
    """world->object 3x4 for an object rotated by `deg` about y and translated by t (inverse = R^T (p - t))."""
    import math
    c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
    # object->world R = [[c,0,s],[0,1,0],[-s,0,c]], world->object = R^T, translation -R^T t
    r = [[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]]
    t = [-(r[i][0] * tx + r[i][1] * ty + r[i][2] * tz) for i in range(3)]
    return tuple(float(np.float32(v)) for i in range(3) for v in (r[i][0], r[i][1], r[i][2], t[i]))


def scene_zoo():
    zoo = {}
    W = H = 96
    # C1-style: ML 48^3, raycast at three sampling rates
    for rate in (0.125, 0.5, 1.0):
        zoo[f"ml48_raycast_r{rate}"] = (default_scene(48, W, H, rate=rate), 1, False)
    # default renderer: jittered pixels, 2 spp, 3 accumulated frames, float colour + aux channels
    s = default_scene(40, W, H, rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT, num_iterations=2,
                      fmt=capi.DVR_FORMAT_FLOAT32_VEC4, channels=("depth", "primId", "objId", "instId", "albedo", "normal"))
    zoo["ml40_default_spp2_f3_float"] = (s, 3, False)
    # UFIXED8_VEC4 (linear) colour, no depth channel
    s = default_scene(32, W, H, rate=0.5, fmt=capi.DVR_FORMAT_UFIXED8_VEC4, channels=())
    zoo["ml32_unorm8_nodepth"] = (s, 1, False)
    # checkerboard: 6 passes (crosses the frameID increment), odd image size
    s = default_scene(32, 75, 53, rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT)
    zoo["ml32_checkerboard_p6"] = (s, 6, True)
    # semi-transparent blobs, large unit distance (long marches, no early termination)
    s = default_scene(48, W, H, rate=1.0, field="blobs")
    s.volumes[0].unit_distance = 1.5
    zoo["blobs48_translucent"] = (s, 1, False)
    # orthographic camera, non-cubic dims, anisotropic spacing, value range outside [0,1], nearest filter variant
    rng = np.random.default_rng(7)
    vox = (rng.random((20, 28, 36), dtype=np.float32) * 3.0 - 1.0).astype(np.float32)
    tf = capi.tf_discretize(color=scenes.tsd_default_colormap(16), opacity=np.array([0.0, 0.2, 0.9, 0.1, 1.0], np.float32),
                            value_range=(-1.0, 2.0))
    for nearest in (False, True):
        v = VolumeDesc(vox, origin=(0.5, -1.0, 2.0), spacing=(0.1, 0.15, 0.2), nearest=nearest, tf=tf,
                       value_range=(-1.0, 2.0), unit_distance=0.4, vol_id=11, inst_id=5)
        lo, hi = v.bounds()
        c = 0.5 * (lo + hi)
        cam = capi.camera_orthographic((float(c[0]) + 3.0, float(c[1]) + 2.0, float(c[2]) + 6.0), (-3.0, -2.0, -6.0),
                                       (0.0, 1.0, 0.0), 6.0, 1.25)
        zoo["noise_ortho_" + ("nearest" if nearest else "linear")] = (
            SceneDesc([v], 100, 80, cam, volume_sampling_rate=0.5, background=(0.0, 0.2, 0.4, 0.5)), 1, False)
    # fixed-point fields
    base = scenes.blobs_np(32)
    for name, dt, arr in (("u8", capi.DVR_UFIXED8, np.round(base * 255).astype(np.uint8)),
                          ("u16", capi.DVR_UFIXED16, np.round(base * 65535).astype(np.uint16)),
                          ("i16", capi.DVR_FIXED16, np.round((base * 2 - 1) * 32767).astype(np.int16)),
                          ("i8", capi.DVR_FIXED8, np.round((base * 2 - 1) * 127).astype(np.int8))):
        s = default_scene(32, W, H, rate=0.5)
        s.volumes[0].voxels = arr
        s.volumes[0].data_type = dt
        s.volumes[0].unit_distance = 0.5
        if name.startswith("i"):
            s.volumes[0].value_range = (-1.0, 1.0)
            s.volumes[0].tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256), value_range=(-1.0, 1.0))
        zoo[f"blobs32_{name}"] = (s, 1, False)
    # two volumes: one rotated+translated instance overlapping the other; thin-lens perspective camera
    a = default_scene(32, W, H, rate=0.5)
    v0 = a.volumes[0]
    v1 = VolumeDesc(scenes.blobs_np(24), origin=(-0.5, -0.5, -0.5), spacing=(1.0 / 23,) * 3,
                    tf=capi.tf_discretize(uniform_color=(0.9, 0.8, 0.1, 0.7), uniform_opacity=0.6),
                    unit_distance=0.2, vol_id=21, inst_id=9, world_to_object=_rot_y(25.0, 0.9, 0.2, 0.4))
    pose = scenes.orbit_camera((-1, -1, -1), (1.5, 1, 1), W, H, az_deg=35.0, el_deg=15.0, dist_scale=1.2)
    cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect, 4.0, 0.05)
    zoo["two_volumes_lens"] = (SceneDesc([v0, v1], W, H, cam, volume_sampling_rate=0.5,
                                         integrator=capi.DVR_INTEGRATOR_DEFAULT), 2, False)
    # camera inside the volume + image region crop
    s = default_scene(32, W, H, rate=0.5)
    s.camera = capi.camera_perspective((0.1, 0.0, 0.2), (0.3, -0.2, -1.0), (0.0, 1.0, 0.0), 1.2, 1.0,
                                       region=(0.1, 0.2, 0.8, 0.9))
    s.volumes[0].unit_distance = 1.0
    zoo["camera_inside_region"] = (s, 1, False)
    # NanoVDB fog spheres (BASELINE config C5, small): committed fixture generated with the reference's NanoVDB
    import os
    fog = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvdb_fog_spheres.npz"))
    for key, ud, rate in (("r20", 4.0, 0.5), ("r12_vs025", 1.0, 1.0)):
        v = VolumeDesc(np.zeros((1, 1, 1), np.float32), nvdb=fog[key], tf=capi.tf_discretize(
            color=scenes.tsd_default_colormap(256)), unit_distance=ud, vol_id=31, inst_id=2)
        lo, hi = v.bounds()
        pose = scenes.orbit_camera(lo, hi, W, H, dist_scale=1.0)
        cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
        zoo[f"nvdb_fog_{key}"] = (SceneDesc([v], W, H, cam, volume_sampling_rate=rate,
                                            integrator=capi.DVR_INTEGRATOR_DEFAULT), 2, False)
    # quantised NanoVDB grids (Fp4 / Fp8 / Fp16 / FpN), quantised by NanoVDB itself (make_nvdb_fixtures.py)
    quant = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nvdb_quant_spheres.npz"))
    for key in ("fp4", "fp8", "fp16", "fpn"):
        v = VolumeDesc(np.zeros((1, 1, 1), np.float32), nvdb=quant[key], tf=capi.tf_discretize(
            color=scenes.tsd_default_colormap(256)), unit_distance=3.0, vol_id=32, inst_id=4)
        lo, hi = v.bounds()
        pose = scenes.orbit_camera(lo, hi, W, H, az_deg=50.0, dist_scale=0.9)
        cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
        zoo[f"nvdb_{key}_r14"] = (SceneDesc([v], W, H, cam, volume_sampling_rate=0.5,
                                            integrator=capi.DVR_INTEGRATOR_DEFAULT), 2, False)
    # renderer "background" as an Array2D image (a18): float RGB (padded to RGBA, alpha 255) behind a translucent volume
    # with the jittered default renderer and 2 spp — every pixel-sample fetches the image at its own jittered screen
    # coordinate; sRGB RGBA (alpha linearised too, the reference's quirk) with the centred raycast renderer and a
    # small volume, so most pixels take the missed-ray path; a 2-channel 16-bit image (texture reads (r, g, 0, 1))
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:13, 0:17]
    img_f = np.stack([xx / 16.0, yy / 12.0, 0.5 + 0.5 * np.sin(xx * 0.9) * np.cos(yy * 0.7)], axis=-1).astype(np.float32)
    s = default_scene(32, W, H, rate=0.5, field="blobs", integrator=capi.DVR_INTEGRATOR_DEFAULT, num_iterations=2,
                      fmt=capi.DVR_FORMAT_FLOAT32_VEC4, channels=("depth", "objId", "albedo"))
    s.volumes[0].unit_distance = 0.6
    s.background_image = (img_f, capi.DVR_IMAGE_FLOAT32)
    zoo["bgimage_float3_default_spp2_f2"] = (s, 2, False)
    img_s = rng.integers(0, 256, (9, 21, 4), dtype=np.uint8)
    s = default_scene(24, 120, 72, rate=0.5)
    pose = scenes.orbit_camera((-1, -1, -1), (1, 1, 1), 120, 72, dist_scale=2.5)
    s.camera = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    s.background_image = (img_s, capi.DVR_IMAGE_SRGB8)
    zoo["bgimage_srgb4_raycast_far"] = (s, 1, False)
    img_u = rng.integers(0, 65536, (6, 5, 2), dtype=np.uint16)
    s = default_scene(24, 64, 48, rate=0.5, integrator=capi.DVR_INTEGRATOR_DEFAULT)
    s.background_image = (img_u, capi.DVR_IMAGE_UFIXED16)
    zoo["bgimage_u16x2_checkerboard_p5"] = (s, 5, True)
    # empty world (no volume instance): background only
    s = default_scene(8, 40, 24, rate=0.5)
    s.volumes = []
    zoo["empty_world"] = (s, 1, False)
    return zoo


# ------------------------------------------------------------------------------------------------------------
# scenes for the dpt renderer (delta tracking); golden fixture: tests/golden/refgpu_dpt.npz
# ------------------------------------------------------------------------------------------------------------
DPT_KINDS = ("ml40", "ml77", "nvdb", "two")
DPT_GOLDEN_FRAMES = 8


# ------------------------------------------------------------------------------------------------------------
# mixed scenes: surfaces in front of / behind / inside volumes, lights, shadow rays
def _quad(p0, p1, p2, p3):
    """two triangles (soup) over the corners p0..p3 (counter-clockwise seen from the front)"""
    return np.array([p0, p1, p2, p0, p2, p3], np.float32)


def _uv_sphere_mesh(center, radius, nu=12, nv=8):
    """indexed triangle mesh of a sphere with per-vertex normals"""
    verts, normals = [], []
    for j in range(nv + 1):
        th = np.pi * j / nv
        for i in range(nu):
            ph = 2.0 * np.pi * i / nu
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            normals.append(n)
            verts.append(np.asarray(center) + radius * n)
    idx = []
    for j in range(nv):
        for i in range(nu):
            a, b = j * nu + i, j * nu + (i + 1) % nu
            c, d = a + nu, b + nu
            idx += [(a, b, d), (a, d, c)]
    return np.array(verts, np.float32), np.array(normals, np.float32), np.array(idx, np.uint32)


def mixed_scene_zoo(wh=96):
    """name -> SceneDesc: the surface branch of the default / raycast renderers around a 32^3 volume on [-1,1]^3"""
    zoo = {}
    sdir = np.array((-0.3, -1.0, -0.2), np.float32)  # normalised as Directional::commitParameters does (fp32)
    sdir = sdir * (np.float32(1.0) / np.sqrt(np.float32(sdir[0] * sdir[0] + sdir[1] * sdir[1] + sdir[2] * sdir[2])))
    sun = {"type": "directional", "direction": tuple(float(c) for c in sdir), "irradiance": 2.5,
           "color": (1.0, 0.95, 0.9)}
    lamp = {"type": "point", "position": (1.6, 1.8, 1.2), "intensity": 6.0}
    floor = {"geometry": "triangle", "vertex.position": _quad((-3, -1.3, -3), (-3, -1.3, 3), (3, -1.3, 3), (3, -1.3, -3)),
             "color": (0.7, 0.7, 0.75), "id": 11, "instanceId": 21}
    balls = {"geometry": "sphere", "vertex.position": [(-0.4, 0.2, 1.3), (0.7, -0.2, 0.2), (0.1, 0.9, -0.6)],
             "vertex.radius": [0.25, 0.35, 0.2], "color": (0.9, 0.3, 0.2), "id": 12, "instanceId": 22,
             "primitive.id": [100, 101, 102]}

    def base(**kw):
        sc = default_scene(32, wh, wh, rate=0.5, field="blobs", integrator=capi.DVR_INTEGRATOR_DEFAULT, **kw)
        sc.volumes[0].unit_distance = 0.25
        return sc

    # shadows of the volume and of the spheres on the floor; spheres in front of, inside and behind the volume
    zoo["floor_balls_sun"] = base(surfaces=[floor, balls], lights=[sun], surface_ambient_radiance=0.2)
    zoo["floor_balls_lamp_spp2"] = base(surfaces=[floor, balls], lights=[lamp, sun], num_iterations=2,
                                        ambient_samples=2, ambient_color=(0.6, 0.7, 1.0), surface_ambient_radiance=0.3)
    # a translucent sheet in front of the volume: the loop continues behind it (blend), a masked one does not exist
    sheet = {"geometry": "triangle", "vertex.position": _quad((-1.2, -1.2, 1.6), (1.2, -1.2, 1.6), (1.2, 1.2, 1.6), (-1.2, 1.2, 1.6)),
             "color": (0.2, 0.6, 0.9, 0.8), "opacity": 0.5, "alphaMode": "blend", "id": 13}
    masked = {"geometry": "triangle", "vertex.position": _quad((-3, -3, 2.2), (3, -3, 2.2), (3, 3, 2.2), (-3, 3, 2.2)),
              "color": (1, 0, 0, 0.3), "alphaMode": "mask", "alphaCutoff": 0.5, "id": 14}
    zoo["translucent_sheet"] = base(surfaces=[sheet, masked, floor], lights=[sun], ambient_samples=0,
                                    surface_ambient_radiance=0.5)
    # raycast renderer: indexed mesh with vertex normals under an instance transform, back-face culling
    v, n, idx = _uv_sphere_mesh((0.0, 0.0, 0.0), 0.5)
    mesh = {"geometry": "triangle", "vertex.position": v, "vertex.normal": n, "primitive.index": idx,
            "color": (0.3, 0.8, 0.4), "id": 15, "instanceId": 25, "cullBackfaces": True,
            "transform": (1.2, 0.0, 0.3, 0.5, 0.0, 0.8, 0.0, -0.3, -0.3, 0.0, 1.2, 1.0)}
    zoo["raycast_mesh_instance"] = base(surfaces=[mesh, floor], ambient_color=(1.0, 0.9, 0.8))
    zoo["raycast_mesh_instance"].integrator = capi.DVR_INTEGRATOR_RAYCAST
    zoo["default_mesh_cullbf"] = base(surfaces=[mesh, balls], lights=[lamp], cull_triangle_backfaces=True,
                                      channels=("depth", "primId", "objId", "instId", "albedo", "normal"))
    # surfaces only (no volume at all) and a float frame
    only = base(surfaces=[floor, balls, mesh], lights=[sun, lamp], surface_ambient_radiance=0.1,
                fmt=capi.DVR_FORMAT_FLOAT32_VEC4)
    only.volumes = []
    zoo["surfaces_only_float"] = only
    return zoo


def dpt_scene(kind, wh=96, **kw) -> SceneDesc:
    if kind == "nvdb":
        from visrtx_b200 import nvdb_writer
        blob = nvdb_writer.fog_sphere(30.0)
        tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256),
                                opacity=np.linspace(0.0, 0.6, 16, dtype=np.float32))
        v = VolumeDesc(np.zeros((1, 1, 1), np.float32), tf=tf, nvdb=blob, unit_distance=1.0)
        lo, hi = v.bounds()
        pose = scenes.orbit_camera(lo, hi, wh, wh, dist_scale=1.2)
        cam = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
        return SceneDesc([v], wh, wh, cam, fmt=capi.DVR_FORMAT_FLOAT32_VEC4, integrator=capi.DVR_INTEGRATOR_DPT,
                         channels=("depth", "objId", "albedo", "normal"), **kw)
    if kind == "two":
        s, _, _ = scene_zoo()["two_volumes_lens"]
        s.integrator = capi.DVR_INTEGRATOR_DPT
        s.num_iterations = 1
        s.fmt = capi.DVR_FORMAT_FLOAT32_VEC4
        if wh != 96:
            raise ValueError("the two-volume scene is fixed at 96x96")
        for k, v in kw.items():
            setattr(s, k, v)
        return s
    n = {"ml40": 40, "ml77": 77}[kind]
    s = default_scene(n, wh, wh, integrator=capi.DVR_INTEGRATOR_DPT, fmt=capi.DVR_FORMAT_FLOAT32_VEC4,
                      channels=("depth", "objId", "albedo", "normal"), **kw)
    # a piecewise opacity so that the grid has empty, thin and dense cells
    s.volumes[0].tf = capi.tf_discretize(color=scenes.tsd_default_colormap(256),
                                         opacity=np.array([0.0, 0.05, 0.8, 0.1, 1.0], np.float32))
    lo, hi = s.volumes[0].bounds()
    pose = scenes.orbit_camera(lo, hi, wh, wh, az_deg=40.0, el_deg=25.0, dist_scale=0.8)  # volume fills the frame
    s.camera = capi.camera_perspective(pose.position, pose.direction, pose.up, pose.fovy, pose.aspect)
    return s


def frac_close(a, b, fmt, tol=2.0):
    """fraction of pixels whose colour differs by <= tol/255"""
    if fmt == capi.DVR_FORMAT_FLOAT32_VEC4:
        d = np.abs(a - b).max(axis=-1) * 255.0
    else:
        d = np.abs(unpack_rgba8(a) - unpack_rgba8(b)).max(axis=-1)
    return float((d <= tol).mean())
