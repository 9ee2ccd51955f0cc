"""ctypes bindings of the ORACLE libraries (test infrastructure; never imported by the product).

  oracle/liboracle_dvr.so        O-cpu   — CPU restatement of the reference path
  oracle/_ref/libref_gpu_dvr.so  O-gpu   — the reference's own device headers compiled for sm_100a
  oracle/_ref/libref_host.so     the reference's own host helpers (TF discretisation)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from visrtx_b200 import pods as capi  # POD structs only: the oracle bindings never load libdvr_b200.so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
CPU_LIB = os.path.join(ORACLE_DIR, "liboracle_dvr.so")
REF_GPU_LIB = os.path.join(ORACLE_DIR, "_ref", "libref_gpu_dvr.so")
REF_HOST_LIB = os.path.join(ORACLE_DIR, "_ref", "libref_host.so")


class OracleVolume(C.Structure):
    _fields_ = [("voxels", C.c_void_p), ("dims", C.c_int32 * 3), ("origin", C.c_float * 3),
                ("spacing", C.c_float * 3), ("filterNearest", C.c_int32), ("tf", C.c_void_p),
                ("valueRange", C.c_float * 2), ("unitDistance", C.c_float), ("id", C.c_uint32),
                ("worldToObject", C.c_float * 12), ("instanceId", C.c_uint32), ("zOwnBegin", C.c_int32),
                ("zOwnEnd", C.c_int32), ("nvdbGrid", C.c_void_p)]


class OracleBuffers(C.Structure):
    _fields_ = [("colorAccumulation", C.c_void_p), ("outColor", C.c_void_p), ("depth", C.c_void_p),
                ("primId", C.c_void_p), ("objId", C.c_void_p), ("instId", C.c_void_p), ("albedo", C.c_void_p),
                ("normal", C.c_void_p)]


class RefInstance(C.Structure):
    _fields_ = [("volume", C.c_void_p), ("worldToObject", C.c_float * 12), ("instanceId", C.c_uint32),
                ("_pad", C.c_uint32)]


_cpu = None
_refgpu = None
_refhost = None


def cpu():
    global _cpu
    if _cpu is None:
        _cpu = C.CDLL(CPU_LIB)
        _cpu.oracle_tex3d.restype = C.c_float
        _cpu.oracle_nvdb_sample.restype = C.c_float
    return _cpu


def have_ref_gpu() -> bool:
    return os.path.exists(REF_GPU_LIB)


def have_ref_host() -> bool:
    return os.path.exists(REF_HOST_LIB)


def refgpu():
    global _refgpu
    if _refgpu is None:
        _refgpu = C.CDLL(REF_GPU_LIB)
        _refgpu.refgpu_last_error.restype = C.c_char_p
    return _refgpu


def refhost():
    global _refhost
    if _refhost is None:
        _refhost = C.CDLL(REF_HOST_LIB)
    return _refhost


def _tf_args(color, opacity, uniform_color, uniform_opacity, value_range):
    cptr, ncol, nch = None, 0, 4
    keep = []
    if color is not None:
        color = np.ascontiguousarray(color, dtype=np.float32)
        keep.append(color)
        ncol, nch = color.shape
        cptr = color.ctypes.data_as(C.c_void_p)
    optr, nop = None, 0
    if opacity is not None:
        opacity = np.ascontiguousarray(opacity, dtype=np.float32)
        keep.append(opacity)
        nop = opacity.shape[0]
        optr = opacity.ctypes.data_as(C.c_void_p)
    return (cptr, C.c_size_t(ncol), C.c_int(nch), optr, C.c_size_t(nop), (C.c_float * 4)(*uniform_color),
            C.c_float(uniform_opacity), (C.c_float * 2)(*value_range)), keep


def tf_discretize(color=None, opacity=None, uniform_color=(1, 1, 1, 1), uniform_opacity=1.0, value_range=(0, 1),
                  which="cpu"):
    out = np.empty((256, 4), dtype=np.float32)
    args, keep = _tf_args(color, opacity, uniform_color, uniform_opacity, value_range)
    fn = cpu().oracle_tf_discretize if which == "cpu" else refhost().refhost_tf_discretize
    rc = fn(*args, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def camera_perspective(pos, direction, up, fovy, aspect, focus_distance=1.0, aperture_radius=0.0, region=None):
    cam = capi.DvrCamera()
    reg = (C.c_float * 4)(*region) if region is not None else None
    cpu().oracle_camera_perspective((C.c_float * 3)(*pos), (C.c_float * 3)(*direction), (C.c_float * 3)(*up),
                                    C.c_float(fovy), C.c_float(aspect), C.c_float(focus_distance),
                                    C.c_float(aperture_radius), reg, C.byref(cam))
    return cam


def camera_orthographic(pos, direction, up, height, aspect, region=None):
    cam = capi.DvrCamera()
    reg = (C.c_float * 4)(*region) if region is not None else None
    cpu().oracle_camera_orthographic((C.c_float * 3)(*pos), (C.c_float * 3)(*direction), (C.c_float * 3)(*up),
                                     C.c_float(height), C.c_float(aspect), reg, C.byref(cam))
    return cam


def tex3d(vol: np.ndarray, u, v, w) -> np.ndarray:
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nz, ny, nx = vol.shape
    dims = (C.c_int * 3)(nx, ny, nz)
    lib = cpu()
    out = np.empty(len(u), dtype=np.float32)
    p = vol.ctypes.data_as(C.c_void_p)
    for i in range(len(u)):
        out[i] = lib.oracle_tex3d(p, dims, C.c_float(u[i]), C.c_float(v[i]), C.c_float(w[i]))
    return out


def tex1d_tf(tf: np.ndarray, coords) -> np.ndarray:
    tf = np.ascontiguousarray(tf, dtype=np.float32)
    out = np.empty((len(coords), 4), dtype=np.float32)
    tmp = (C.c_float * 4)()
    for i, c in enumerate(coords):
        cpu().oracle_tex1d_tf(tf.ctypes.data_as(C.c_void_p), C.c_float(c), tmp)
        out[i] = tuple(tmp)
    return out


def philox_uniforms(seed: int, offset: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    cpu().oracle_philox_uniforms(C.c_uint64(seed), C.c_uint64(offset), C.c_int(n), out.ctypes.data_as(C.c_void_p))
    return out


def nvdb_fog_sphere(radius=20.0, voxel_size=1.0, half_width=3.0, center=(0.0, 0.0, 0.0)) -> np.ndarray:
    """A serialized NanoVDB float fog-sphere grid from the reference's vendored NanoVDB (needs libref_host.so)."""
    lib = refhost()
    lib.refhost_nvdb_fog_sphere.restype = C.c_size_t
    ctr = (C.c_double * 3)(*center)
    n = lib.refhost_nvdb_fog_sphere(C.c_double(radius), C.c_double(voxel_size), C.c_double(half_width), ctr, None,
                                    C.c_size_t(0))
    buf = np.zeros(n, np.uint8)
    lib.refhost_nvdb_fog_sphere(C.c_double(radius), C.c_double(voxel_size), C.c_double(half_width), ctr,
                                buf.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    return buf


NVDB_GRID_TYPES = {"float": 1, "fp4": 13, "fp8": 14, "fp16": 15, "fpn": 16}  # nanovdb::GridType


def nvdb_fog_sphere_typed(grid_type="fp8", radius=20.0, voxel_size=1.0, half_width=3.0, center=(0.0, 0.0, 0.0),
                          tolerance=-1.0) -> np.ndarray:
    """Same sphere through NanoVDB's own quantiser: Fp4 / Fp8 / Fp16 / FpN grids (needs libref_host.so)."""
    lib = refhost()
    lib.refhost_nvdb_fog_sphere_typed.restype = C.c_size_t
    ctr = (C.c_double * 3)(*center)
    args = (C.c_uint(NVDB_GRID_TYPES[grid_type]), C.c_double(radius), C.c_double(voxel_size), C.c_double(half_width),
            ctr, C.c_float(tolerance))
    n = lib.refhost_nvdb_fog_sphere_typed(*args, None, C.c_size_t(0))
    assert n > 0
    buf = np.zeros(n, np.uint8)
    lib.refhost_nvdb_fog_sphere_typed(*args, buf.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    return buf


def nvdb_is_valid_reference(blob: np.ndarray) -> bool:
    return bool(refhost().refhost_nvdb_is_valid(blob.ctypes.data_as(C.c_void_p)))


def nvdb_sample_reference(blob: np.ndarray, xyz: np.ndarray) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, np.float32)
    out = np.empty(len(xyz), np.float32)
    refhost().refhost_nvdb_sample(blob.ctypes.data_as(C.c_void_p), xyz.ctypes.data_as(C.c_void_p), C.c_int(len(xyz)),
                                  out.ctypes.data_as(C.c_void_p))
    return out


def nvdb_sample_oracle(blob: np.ndarray, xyz: np.ndarray) -> np.ndarray:
    lib = cpu()
    p = blob.ctypes.data_as(C.c_void_p)
    return np.array([lib.oracle_nvdb_sample(p, C.c_float(a), C.c_float(b), C.c_float(c)) for a, b, c in xyz], np.float32)
