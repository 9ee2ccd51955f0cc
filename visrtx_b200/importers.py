"""ctypes binding of include/dvr_import.h (libdvr_import.so) + the ANARI-side convenience of the reference's
``tsd::import_volume`` (tsd/src/tsd/authoring/importers/import_volume.cpp:12-66): a volume file becomes a
spatial field (``structuredRegular`` or ``nanovdb``) bound to a ``transferFunction1D`` volume with the TSD
default colour map and ``valueRange`` = the data range."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_LIB_PATH = os.environ.get("DVR_IMPORT_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdvr_import.so"))
if not os.path.exists(_LIB_PATH):
    raise ImportError(f"{_LIB_PATH} is missing: build it with `make -C visrtx_b200/importers`")
lib = C.CDLL(_LIB_PATH)

EXPORTED_SYMBOLS = ["dvr_import_last_error", "dvr_import_raw", "dvr_import_mhd", "dvr_import_vti", "dvr_import_nvdb",
                    "dvr_import_volume", "dvr_import_free", "dvr_compute_scalar_range"]

OK, ERR_ARGUMENT, ERR_IO, ERR_FORMAT, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
STRUCTURED, NANOVDB = 0, 1
# DvrDataType -> numpy
NP_TYPES = {0: np.float32, 1: np.uint8, 2: np.int8, 3: np.uint16, 4: np.int16, 5: np.float64, 6: np.float16}


class DvrVolumeFile(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dataType", C.c_int32), ("dims", C.c_uint32 * 3), ("origin", C.c_float * 3),
                ("spacing", C.c_float * 3), ("headerSpacing", C.c_double * 3), ("valueRange", C.c_float * 2),
                ("hasValueRange", C.c_int32), ("_pad", C.c_int32), ("data", C.c_void_p), ("bytes", C.c_uint64),
                ("name", C.c_char * 256)]


lib.dvr_import_last_error.restype = C.c_char_p
for _n in ("dvr_import_raw", "dvr_import_mhd", "dvr_import_vti", "dvr_import_nvdb", "dvr_import_volume"):
    getattr(lib, _n).argtypes = [C.c_char_p, C.POINTER(DvrVolumeFile)]
lib.dvr_import_free.argtypes = [C.POINTER(DvrVolumeFile)]
lib.dvr_compute_scalar_range.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(C.c_float)]


class ImportError_(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


@dataclass
class VolumeFile:
    kind: int
    data_type: int
    dims: tuple
    origin: tuple
    spacing: tuple
    header_spacing: tuple
    value_range: tuple
    has_value_range: bool
    data: np.ndarray  # structured: (z,y,x) voxels; nanovdb: uint8 blob
    name: str


def _import(fn, path: str) -> VolumeFile:
    f = DvrVolumeFile()
    rc = fn(os.fsencode(path), C.byref(f))
    if rc != OK:
        raise ImportError_(rc, lib.dvr_import_last_error().decode())
    try:
        raw = np.ctypeslib.as_array(C.cast(f.data, C.POINTER(C.c_uint8)), shape=(int(f.bytes),)).copy()
        if f.kind == STRUCTURED:
            nx, ny, nz = f.dims
            data = raw.view(NP_TYPES[f.dataType]).reshape(nz, ny, nx)
        else:
            data = raw
        return VolumeFile(f.kind, f.dataType, tuple(f.dims), tuple(f.origin), tuple(f.spacing), tuple(f.headerSpacing),
                          tuple(f.valueRange), bool(f.hasValueRange), data, f.name.decode())
    finally:
        lib.dvr_import_free(C.byref(f))


def import_raw(path): return _import(lib.dvr_import_raw, path)
def import_mhd(path): return _import(lib.dvr_import_mhd, path)
def import_vti(path): return _import(lib.dvr_import_vti, path)
def import_nvdb(path): return _import(lib.dvr_import_nvdb, path)
def import_volume_file(path): return _import(lib.dvr_import_volume, path)


def compute_scalar_range(arr: np.ndarray, data_type: int):
    a = np.ascontiguousarray(arr)
    out = (C.c_float * 2)()
    rc = lib.dvr_compute_scalar_range(a.ctypes.data_as(C.c_void_p), data_type, a.size, out)
    if rc != OK:
        raise ImportError_(rc, lib.dvr_import_last_error().decode())
    return float(out[0]), float(out[1])


def import_volume(device, path: str, color=None, opacity=None):
    """tsd::import_volume on an ANARI device (visrtx_b200.anari.Device): returns (volume, field, VolumeFile).
    The caller owns both handles (release them when done)."""
    from . import anari as A
    from . import scenes
    vf = import_volume_file(path)
    d = device
    if vf.kind == STRUCTURED:
        elem = {0: A.FLOAT32, 1: A.UFIXED8, 2: A.FIXED8, 3: A.UFIXED16, 4: A.FIXED16, 5: A.FLOAT64}[vf.data_type]
        arr = d.new_array3d(vf.data, elem)
        field = d.new("SpatialField", "structuredRegular")
        d.set(field, "data", A.ARRAY3D, arr)
        d.set(field, "origin", A.FLOAT32_VEC3, vf.origin)
        d.set(field, "spacing", A.FLOAT32_VEC3, vf.spacing)
    else:
        arr = d.new_array1d(vf.data, A.UINT8)
        field = d.new("SpatialField", "nanovdb")
        d.set(field, "data", A.ARRAY1D, arr)
    d.commit(field)
    d.release(arr)
    volume = d.new("Volume", "transferFunction1D")
    cmap = scenes.tsd_default_colormap(256) if color is None else np.asarray(color, np.float32)
    carr = d.new_array1d(cmap, A.FLOAT32_VEC4 if cmap.shape[1] == 4 else A.FLOAT32_VEC3)
    d.set(volume, "color", A.ARRAY1D, carr)
    d.release(carr)
    if opacity is not None:
        oarr = d.new_array1d(np.asarray(opacity, np.float32), A.FLOAT32)
        d.set(volume, "opacity", A.ARRAY1D, oarr)
        d.release(oarr)
    d.set(volume, "value", A.SPATIAL_FIELD, field)
    d.set(volume, "valueRange", A.FLOAT32_BOX1, vf.value_range)
    d.commit(volume)
    return volume, field, vf
