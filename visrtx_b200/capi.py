"""ctypes binding of the C-ABI launch layer (include/dvr_b200.h).

One Python function per exported symbol, same names and argument meaning.  Errors are raised as
:class:`DvrError` carrying ``dvr_last_error()``; nothing here computes anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DVR_B200_LIB") or os.path.join(_HERE, "libdvr_b200.so")

from .pods import *  # noqa: F401,F403  (enums, POD structs, frame_params / frame_buffers / peer_sync)
from .pods import DvrCamera, DvrVolumeInstance, DvrFrameBuffers, DvrFrameParams, DvrRenderStats, DvrPeerSync, DvrSlabExchange
from .pods import DvrSurfaceDesc, DvrLight, DvrSceneParams, surface_descs, scene_params

# every symbol include/dvr_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "dvr_last_error", "dvr_version", "dvr_device_count", "dvr_set_device", "dvr_device_info",
    "dvr_camera_perspective", "dvr_camera_orthographic", "dvr_tf_discretize",
    "dvr_field_create_structured", "dvr_field_create_structured_slab", "dvr_field_upload_slices", "dvr_field_set_owned_slices", "dvr_field_owned_slices", "dvr_field_update_structured", "dvr_field_create_nanovdb",
    "dvr_field_destroy",
    "dvr_field_bounds", "dvr_field_step_size", "dvr_field_device_bytes", "dvr_field_build_macrocells",
    "dvr_field_macrocells", "dvr_field_value_range",
    "dvr_volume_create", "dvr_volume_update", "dvr_volume_destroy", "dvr_volume_majorants",
    "dvr_volume_dda_majorants", "dvr_image_create", "dvr_image_destroy",
    "dvr_post_convert_float_color", "dvr_post_composite_depth", "dvr_post_outline", "dvr_post_visualize_depth",
    "dvr_post_pick", "dvr_selftest_lattice_advance", "dvr_bounds_screen_rect",
    "dvr_render", "dvr_render_instrumented", "dvr_launch_count",
    "dvr_surfaces_create", "dvr_surfaces_destroy", "dvr_surfaces_info", "dvr_render_scene",
    "dvr_render_partial", "dvr_render_partial_instrumented", "dvr_composite_over", "dvr_resolve", "dvr_scale_vec3",
    "dvr_composite_resolve_peers", "dvr_render_partial_sync", "dvr_composite_resolve_peers_sync", "dvr_wait_flags",
    "dvr_render_slab_frame",
    "dvr_ipc_alloc", "dvr_ipc_open", "dvr_ipc_close", "dvr_ipc_free",
]


class DvrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dvr error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C visrtx_b200/csrc`).  There is no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()
lib.dvr_last_error.restype = C.c_char_p
lib.dvr_launch_count.restype = C.c_ulonglong
for _name in EXPORTED_SYMBOLS:
    if _name not in ("dvr_last_error", "dvr_launch_count"):
        getattr(lib, _name).restype = C.c_int

_f3 = C.c_float * 3
_f4 = C.c_float * 4
_f2 = C.c_float * 2
_u3 = C.c_uint32 * 3


def _check(rc: int) -> None:
    if rc != DVR_OK:
        raise DvrError(rc, lib.dvr_last_error().decode())


def last_error() -> str:
    return lib.dvr_last_error().decode()


def version():
    a, b = C.c_int(), C.c_int()
    _check(lib.dvr_version(C.byref(a), C.byref(b)))
    return a.value, b.value


def device_count() -> int:
    return int(lib.dvr_device_count())


def set_device(i: int) -> None:
    _check(lib.dvr_set_device(C.c_int(i)))


def device_info():
    name = C.create_string_buffer(256)
    sms, mem = C.c_int(), C.c_size_t()
    _check(lib.dvr_device_info(name, C.c_size_t(256), C.byref(sms), C.byref(mem)))
    return name.value.decode(), sms.value, mem.value


def launch_count() -> int:
    return int(lib.dvr_launch_count())


def camera_perspective(pos, direction, up, fovy, aspect, focus_distance=1.0, aperture_radius=0.0,
                       region=None) -> DvrCamera:
    cam = DvrCamera()
    reg = _f4(*region) if region is not None else None
    _check(lib.dvr_camera_perspective(_f3(*pos), _f3(*direction), _f3(*up), C.c_float(fovy), C.c_float(aspect),
                                      C.c_float(focus_distance), C.c_float(aperture_radius), reg, C.byref(cam)))
    return cam


def camera_orthographic(pos, direction, up, height, aspect, region=None) -> DvrCamera:
    cam = DvrCamera()
    reg = _f4(*region) if region is not None else None
    _check(lib.dvr_camera_orthographic(_f3(*pos), _f3(*direction), _f3(*up), C.c_float(height), C.c_float(aspect),
                                       reg, C.byref(cam)))
    return cam


def tf_discretize(color=None, opacity=None, uniform_color=(1.0, 1.0, 1.0, 1.0), uniform_opacity=1.0,
                  value_range=(0.0, 1.0)):
    """TransferFunction1D::discritizeTFData; color is an (N,3|4) float32 numpy array or None."""
    import numpy as np
    out = np.empty((DVR_TF_SIZE, 4), dtype=np.float32)
    cptr, ncol, nch = None, 0, 4
    if color is not None:
        color = np.ascontiguousarray(color, dtype=np.float32)
        ncol, nch = color.shape
        cptr = color.ctypes.data_as(C.c_void_p)
    optr, nop = None, 0
    if opacity is not None:
        opacity = np.ascontiguousarray(opacity, dtype=np.float32)
        nop = opacity.shape[0]
        optr = opacity.ctypes.data_as(C.c_void_p)
    _check(lib.dvr_tf_discretize(cptr, C.c_size_t(ncol), C.c_int(nch), optr, C.c_size_t(nop), _f4(*uniform_color),
                                 C.c_float(uniform_opacity), _f2(*value_range), out.ctypes.data_as(C.c_void_p)))
    return out


class Field:
    """DvrField handle (StructuredRegularField GPU state)."""

    def __init__(self, handle: C.c_void_p):
        self.handle = handle

    @staticmethod
    def create_structured(data_ptr: int, is_device: bool, data_type: int, dims: Sequence[int], origin, spacing,
                          filter_mode: int = DVR_FILTER_LINEAR, stream: int = 0) -> "Field":
        h = C.c_void_p()
        _check(lib.dvr_field_create_structured(C.c_void_p(data_ptr), C.c_int(1 if is_device else 0),
                                               C.c_int(data_type), _u3(*dims), _f3(*origin), _f3(*spacing),
                                               C.c_int(filter_mode), C.c_void_p(stream), C.byref(h)))
        return Field(h)

    @staticmethod
    def create_slab(data_ptr: int, is_device: bool, data_type: int, global_dims, z_begin: int, z_end: int, origin,
                    spacing, filter_mode: int = DVR_FILTER_LINEAR, stream: int = 0) -> "Field":
        h = C.c_void_p()
        _check(lib.dvr_field_create_structured_slab(C.c_void_p(data_ptr), C.c_int(1 if is_device else 0),
                                                    C.c_int(data_type), _u3(*global_dims), C.c_uint32(z_begin),
                                                    C.c_uint32(z_end), _f3(*origin), _f3(*spacing),
                                                    C.c_int(filter_mode), C.c_void_p(stream), C.byref(h)))
        return Field(h)

    @staticmethod
    def create_nanovdb(data_ptr: int, nbytes: int, is_device: bool = False, stream: int = 0) -> "Field":
        h = C.c_void_p()
        _check(lib.dvr_field_create_nanovdb(C.c_void_p(data_ptr), C.c_size_t(nbytes), C.c_int(1 if is_device else 0),
                                            C.c_void_p(stream), C.byref(h)))
        return Field(h)

    def upload_slices(self, data_ptr: int, is_device: bool, first_resident_slice: int, n_slices: int,
                      stream: int = 0) -> None:
        _check(lib.dvr_field_upload_slices(self.handle, C.c_void_p(data_ptr), C.c_int(1 if is_device else 0),
                                           C.c_uint32(first_resident_slice), C.c_uint32(n_slices), C.c_void_p(stream)))

    def set_owned_slices(self, z_begin: int, z_end: int) -> None:
        """Sort-last load balancing: move the slab's ownership inside the range it was created with (no data moves)."""
        _check(lib.dvr_field_set_owned_slices(self.handle, C.c_uint32(z_begin), C.c_uint32(z_end)))

    def owned_slices(self):
        """((z_begin, z_end), (limit_begin, limit_end))"""
        v = [C.c_uint32() for _ in range(4)]
        _check(lib.dvr_field_owned_slices(self.handle, *[C.byref(x) for x in v]))
        return (v[0].value, v[1].value), (v[2].value, v[3].value)

    def update_structured(self, data_ptr: int, is_device: bool, data_type: int, origin, spacing,
                          stream: int = 0) -> None:
        """In-place refresh of a whole structured field (same dims / type / filter); volumes bound to it then need
        Volume.update."""
        _check(lib.dvr_field_update_structured(self.handle, C.c_void_p(data_ptr), C.c_int(1 if is_device else 0),
                                               C.c_int(data_type), _f3(*origin), _f3(*spacing), C.c_void_p(stream)))

    def destroy(self) -> None:
        if self.handle:
            lib.dvr_field_destroy(self.handle)
            self.handle = None

    def bounds(self):
        lo, hi = _f3(), _f3()
        _check(lib.dvr_field_bounds(self.handle, lo, hi))
        return tuple(lo), tuple(hi)

    def step_size(self) -> float:
        s = C.c_float()
        _check(lib.dvr_field_step_size(self.handle, C.byref(s)))
        return s.value

    def device_bytes(self) -> int:
        b = C.c_size_t()
        _check(lib.dvr_field_device_bytes(self.handle, C.byref(b)))
        return b.value

    def build_macrocells(self, stream: int = 0) -> None:
        _check(lib.dvr_field_build_macrocells(self.handle, C.c_void_p(stream)))

    def macrocells(self):
        dims = _u3()
        ptr = C.c_void_p()
        _check(lib.dvr_field_macrocells(self.handle, dims, C.byref(ptr)))
        return tuple(dims), ptr.value

    def value_range(self, stream: int = 0):
        r = _f2()
        _check(lib.dvr_field_value_range(self.handle, C.c_void_p(stream), r))
        return r[0], r[1]


class Volume:
    """DvrVolume handle (TransferFunction1D GPU state)."""

    def __init__(self, handle: C.c_void_p, field: Field):
        self.handle = handle
        self.field = field

    @staticmethod
    def create(field: Field, tf_rgba, value_range=(0.0, 1.0), unit_distance=1.0, vol_id=0xFFFFFFFF,
               stream: int = 0) -> "Volume":
        import numpy as np
        tf = np.ascontiguousarray(tf_rgba, dtype=np.float32).reshape(DVR_TF_SIZE, 4)
        h = C.c_void_p()
        _check(lib.dvr_volume_create(field.handle, tf.ctypes.data_as(C.c_void_p), _f2(*value_range),
                                     C.c_float(unit_distance), C.c_uint32(vol_id), C.c_void_p(stream), C.byref(h)))
        return Volume(h, field)

    def update(self, tf_rgba, value_range=(0.0, 1.0), unit_distance=1.0, vol_id=0xFFFFFFFF, stream: int = 0):
        import numpy as np
        tf = np.ascontiguousarray(tf_rgba, dtype=np.float32).reshape(DVR_TF_SIZE, 4)
        _check(lib.dvr_volume_update(self.handle, tf.ctypes.data_as(C.c_void_p), _f2(*value_range),
                                     C.c_float(unit_distance), C.c_uint32(vol_id), C.c_void_p(stream)))

    def majorants_ptr(self) -> int:
        p = C.c_void_p()
        _check(lib.dvr_volume_majorants(self.handle, C.byref(p)))
        return p.value

    def dda_majorants(self, stream: int = 0, reference_build: bool = False):
        """(dims, device pointer) of the delta-tracking grid; builds it if needed.  reference_build: the content
        the reference computes (quirks Q7/Q8) instead of the conservative one."""
        p = C.c_void_p()
        dims = (C.c_uint32 * 3)()
        _check(lib.dvr_volume_dda_majorants(self.handle, C.c_int32(1 if reference_build else 0), C.c_void_p(stream),
                                            dims, C.byref(p)))
        return tuple(dims), p.value

    def destroy(self) -> None:
        if self.handle:
            lib.dvr_volume_destroy(self.handle)
            self.handle = None


def make_instances(volumes: Sequence[Volume], xfms=None, inst_ids=None):
    n = len(volumes)
    arr = (DvrVolumeInstance * max(n, 1))()
    for i, v in enumerate(volumes):
        arr[i].volume = v.handle
        m = IDENTITY_3X4 if xfms is None or xfms[i] is None else tuple(float(x) for x in xfms[i])
        arr[i].worldToObject = (C.c_float * 12)(*m)
        arr[i].instanceId = 0xFFFFFFFF if inst_ids is None else int(inst_ids[i])
    return arr, n




class Image:
    """DvrImage: the renderer's background image (host pixels [h, w, channels] -> RGBA8 texture on the device)."""

    def __init__(self, handle):
        self.handle = handle

    @staticmethod
    def create(pixels, component_type: int, stream: int = 0) -> "Image":
        import numpy as np
        pixels = np.ascontiguousarray(pixels)
        if pixels.ndim == 2:
            pixels = pixels[:, :, None]
        h, w, c = pixels.shape
        out = C.c_void_p()
        _check(lib.dvr_image_create(pixels.ctypes.data_as(C.c_void_p), C.c_int(component_type), C.c_int(c), C.c_uint32(w),
                                    C.c_uint32(h), C.c_void_p(stream), C.byref(out)))
        return Image(out)

    def destroy(self):
        if self.handle:
            _check(lib.dvr_image_destroy(self.handle))
            self.handle = None


def render(params: DvrFrameParams, camera: DvrCamera, instances, n_instances: int, buffers: DvrFrameBuffers,
           stream: int = 0) -> None:
    _check(lib.dvr_render(C.byref(params), C.byref(camera), instances, C.c_uint32(n_instances), C.byref(buffers),
                          C.c_void_p(stream)))


class Surfaces:
    """DvrSurfaces handle: the flattened world's surfaces (geometry + matte material + instance transform) and their BVH."""

    def __init__(self, handle):
        self.handle = handle

    @staticmethod
    def create(surfaces, stream: int = 0) -> "Surfaces":
        arr, keep = surface_descs(surfaces)
        h = C.c_void_p()
        _check(lib.dvr_surfaces_create(arr, C.c_uint32(len(surfaces)), C.c_void_p(stream), C.byref(h)))
        del keep
        return Surfaces(h)

    def info(self):
        n, m = C.c_uint32(), C.c_uint32()
        _check(lib.dvr_surfaces_info(self.handle, C.byref(n), C.byref(m)))
        return n.value, m.value

    def destroy(self):
        if self.handle:
            _check(lib.dvr_surfaces_destroy(self.handle))
            self.handle = None


def render_scene(params, camera, instances, n_instances: int, scene: DvrSceneParams, buffers, stream: int = 0):
    _check(lib.dvr_render_scene(C.byref(params), C.byref(camera), instances, C.c_uint32(n_instances), C.byref(scene),
                                C.byref(buffers), C.c_void_p(stream)))


def render_instrumented(params, camera, instances, n_instances, buffers, stats_dev_ptr: int, stream: int = 0):
    _check(lib.dvr_render_instrumented(C.byref(params), C.byref(camera), instances, C.c_uint32(n_instances),
                                       C.byref(buffers), C.c_void_p(stats_dev_ptr), C.c_void_p(stream)))


def render_partial(params, camera, instance, partial_rgba: int, partial_depth: int, stream: int = 0):
    _check(lib.dvr_render_partial(C.byref(params), C.byref(camera), instance, C.c_void_p(partial_rgba),
                                  C.c_void_p(partial_depth), C.c_void_p(stream)))


def render_partial_instrumented(params, camera, instance, partial_rgba: int, partial_depth: int, stats_dev_ptr: int,
                                stream: int = 0):
    _check(lib.dvr_render_partial_instrumented(C.byref(params), C.byref(camera), instance, C.c_void_p(partial_rgba),
                                               C.c_void_p(partial_depth), C.c_void_p(stats_dev_ptr),
                                               C.c_void_p(stream)))


def composite_over(front_rgba: int, front_depth: int, back_rgba: int, back_depth: int, begin: int, end: int,
                   back_is_in_front: bool, stream: int = 0):
    _check(lib.dvr_composite_over(C.c_void_p(front_rgba), C.c_void_p(front_depth or None), C.c_void_p(back_rgba),
                                  C.c_void_p(back_depth or None), C.c_size_t(begin), C.c_size_t(end),
                                  C.c_int(1 if back_is_in_front else 0), C.c_void_p(stream)))


def resolve(params, partial_rgba: int, partial_depth: int, obj_id: int, inst_id: int, buffers, begin: int, end: int,
            stream: int = 0):
    _check(lib.dvr_resolve(C.byref(params), C.c_void_p(partial_rgba), C.c_void_p(partial_depth or None),
                           C.c_uint32(obj_id), C.c_uint32(inst_id), C.byref(buffers), C.c_size_t(begin),
                           C.c_size_t(end), C.c_void_p(stream)))


def scale_vec3(src: int, dst: int, n_pixels: int, scale: float, stream: int = 0):
    _check(lib.dvr_scale_vec3(C.c_void_p(src), C.c_void_p(dst), C.c_size_t(n_pixels), C.c_float(scale),
                              C.c_void_p(stream)))


def composite_resolve_peers(params, camera, partial_rgba_ptrs, partial_depth_ptrs, obj_id: int, inst_id: int,
                            buffers, begin: int, end: int, stream: int = 0):
    n = len(partial_rgba_ptrs)
    rg = (C.c_void_p * n)(*partial_rgba_ptrs)
    dp = (C.c_void_p * n)(*partial_depth_ptrs) if partial_depth_ptrs else None
    _check(lib.dvr_composite_resolve_peers(C.byref(params), C.byref(camera), rg, dp, C.c_uint32(n),
                                           C.c_uint32(obj_id), C.c_uint32(inst_id), C.byref(buffers),
                                           C.c_size_t(begin), C.c_size_t(end), C.c_void_p(stream)))


def ipc_alloc(nbytes: int):
    """cudaMalloc + cudaIpcGetMemHandle: returns (device pointer, 64-byte handle)."""
    p = C.c_void_p()
    h = (C.c_ubyte * 64)()
    _check(lib.dvr_ipc_alloc(C.c_size_t(nbytes), C.byref(p), h))
    return p.value, bytes(h)


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    h = (C.c_ubyte * 64)(*handle)
    _check(lib.dvr_ipc_open(h, C.byref(p)))
    return p.value


def ipc_close(ptr: int) -> None:
    _check(lib.dvr_ipc_close(C.c_void_p(ptr)))


def ipc_free(ptr: int) -> None:
    _check(lib.dvr_ipc_free(C.c_void_p(ptr)))


def render_partial_sync(params, camera, instance, partial_rgba: int, partial_depth: int, sync: DvrPeerSync,
                        stream: int = 0):
    _check(lib.dvr_render_partial_sync(C.byref(params), C.byref(camera), instance, C.c_void_p(partial_rgba),
                                       C.c_void_p(partial_depth), C.byref(sync), C.c_void_p(stream)))


def composite_resolve_peers_sync(params, camera, instance, partial_rgba_ptrs, partial_depth_ptrs, obj_id: int,
                                 inst_id: int, buffers, begin: int, end: int, sync: DvrPeerSync, stream: int = 0):
    n = len(partial_rgba_ptrs)
    rg = (C.c_void_p * n)(*partial_rgba_ptrs)
    dp = (C.c_void_p * n)(*partial_depth_ptrs) if partial_depth_ptrs else None
    _check(lib.dvr_composite_resolve_peers_sync(C.byref(params), C.byref(camera), instance, rg, dp, C.c_uint32(n),
                                                C.c_uint32(obj_id), C.c_uint32(inst_id), C.byref(buffers),
                                                C.c_size_t(begin), C.c_size_t(end), C.byref(sync), C.c_void_p(stream)))


def render_slab_frame(params, camera, instance, obj_id: int, inst_id: int, buffers, exchange: DvrSlabExchange,
                      stream: int = 0):
    _check(lib.dvr_render_slab_frame(C.byref(params), C.byref(camera), instance, C.c_uint32(obj_id), C.c_uint32(inst_id),
                                     C.byref(buffers), C.byref(exchange), C.c_void_p(stream)))


def wait_flags(flags_ptr: int, n: int, value: int, error_flag: int = 0, stream: int = 0):
    _check(lib.dvr_wait_flags(C.c_void_p(flags_ptr), C.c_uint32(n), C.c_uint32(int(value) & 0xFFFFFFFF),
                              C.c_void_p(error_flag or None), C.c_void_p(stream)))


# ---- frame post passes (tsd/src/render_pipeline/passes) on device pointers ------------------------------------
def post_convert_float_color(rgba_f32_ptr: int, rgba8_ptr: int, n_pixels: int, stream: int = 0):
    _check(lib.dvr_post_convert_float_color(C.c_void_p(rgba_f32_ptr), C.c_void_p(rgba8_ptr), C.c_size_t(n_pixels),
                                            C.c_void_p(stream)))


def post_composite_depth(color_out: int, depth_out: int, id_out: int, color_in: int, depth_in: int, id_in: int,
                         n_pixels: int, first_pass: bool, stream: int = 0):
    _check(lib.dvr_post_composite_depth(C.c_void_p(color_out), C.c_void_p(depth_out), C.c_void_p(id_out or None),
                                        C.c_void_p(color_in), C.c_void_p(depth_in), C.c_void_p(id_in or None),
                                        C.c_size_t(n_pixels), C.c_int(1 if first_pass else 0), C.c_void_p(stream)))


def post_outline(rgba8_ptr: int, object_id_ptr: int, width: int, height: int, outline_id: int, stream: int = 0):
    _check(lib.dvr_post_outline(C.c_void_p(rgba8_ptr), C.c_void_p(object_id_ptr), C.c_uint32(width), C.c_uint32(height),
                                C.c_uint32(outline_id), C.c_void_p(stream)))


def post_visualize_depth(rgba8_ptr: int, depth_ptr: int, n_pixels: int, max_depth: float, stream: int = 0):
    _check(lib.dvr_post_visualize_depth(C.c_void_p(rgba8_ptr), C.c_void_p(depth_ptr), C.c_size_t(n_pixels),
                                        C.c_float(max_depth), C.c_void_p(stream)))


def post_pick(depth_ptr: int, object_id_ptr: int, width: int, height: int, x: int, y: int, stream: int = 0):
    d, i = C.c_float(), C.c_uint32()
    _check(lib.dvr_post_pick(C.c_void_p(depth_ptr), C.c_void_p(object_id_ptr or None), C.c_uint32(width),
                             C.c_uint32(height), C.c_uint32(x), C.c_uint32(y), C.byref(d), C.byref(i), C.c_void_p(stream)))
    return d.value, i.value


def selftest_lattice_advance(count: int = 1 << 20, seed: int = 1, stream: int = 0) -> int:
    """Number of operand sets on which the closed-form lattice advance differs from the literal add loop (0)."""
    m = C.c_uint32(0xFFFFFFFF)
    _check(lib.dvr_selftest_lattice_advance(C.c_uint32(count), C.c_uint64(seed), C.byref(m), C.c_void_p(stream)))
    return m.value


def bounds_screen_rect(camera: DvrCamera, lo, hi, width: int, height: int):
    """(valid, (x0, y0, x1, y1)): the conservative pixel rectangle of an axis-aligned box (host arithmetic only)."""
    r = (C.c_int32 * 4)()
    rc = lib.dvr_bounds_screen_rect(C.byref(camera), _f3(*lo), _f3(*hi), C.c_uint32(width), C.c_uint32(height), r)
    if rc < 0:
        _check(rc)
    return bool(rc), tuple(r)
