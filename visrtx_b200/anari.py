"""ctypes mirror of the ANARI C API exported by ``libanari_library_visrtx_b200.so``.

Thin, allocation-free wrappers with the names of the ANARI C++ convenience API the reference's
applications use (``newObject``, ``setParameter``, ``commitParameters``, ``render``, ``wait``, ``map`` ...;
see examples/simple/testApp_spheres.cpp:167-263 of the reference) so the tests read like ANARI apps.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ANARI_B200_LIB") or os.path.join(_HERE, "libanari_library_visrtx_b200.so")

# enums (include/anari/anari.h)
UNKNOWN = 0
DATA_TYPE, STRING, VOID_POINTER, BOOL = 100, 101, 102, 103
STRING_LIST, DATA_TYPE_LIST, PARAMETER_LIST = 150, 151, 152
STATUS_CALLBACK, FRAME_COMPLETION_CALLBACK = 202, 203
DEVICE, ARRAY1D, ARRAY2D, ARRAY3D, CAMERA, FRAME, GROUP, INSTANCE, RENDERER, SPATIAL_FIELD, VOLUME, WORLD = (
    501, 504, 505, 506, 507, 508, 510, 511, 514, 517, 518, 519)
GEOMETRY, LIGHT, MATERIAL, SURFACE, SAMPLER = 509, 512, 513, 515, 516
UINT8, INT32, UINT32, UINT32_VEC2, UINT32_VEC3, UINT64 = 1004, 1016, 1020, 1021, 1022, 1028
FIXED8, UFIXED8, UFIXED8_VEC4, FIXED16, UFIXED16 = 1032, 1036, 1039, 1040, 1044
UFIXED8_VEC3, UFIXED16_VEC2 = 1038, 1045
FLOAT16, FLOAT32, FLOAT32_VEC2, FLOAT32_VEC3, FLOAT32_VEC4, FLOAT64 = 1064, 1068, 1069, 1070, 1071, 1072
UFIXED8_RGBA_SRGB = 2003
FLOAT32_BOX1, FLOAT32_BOX2, FLOAT32_BOX3 = 2008, 2009, 2010
FLOAT32_MAT4, FLOAT32_MAT3x4 = 2014, 2016
FLOAT64_BOX1 = 2208
SEVERITY_FATAL_ERROR, SEVERITY_ERROR, SEVERITY_WARNING, SEVERITY_PERFORMANCE_WARNING, SEVERITY_INFO, SEVERITY_DEBUG = (
    6000, 6001, 6002, 6003, 6004, 6005)
NO_WAIT, WAIT = 0, 1

API_SYMBOLS = [
    "anariLoadLibrary", "anariUnloadLibrary", "anariLoadModule", "anariUnloadModule", "anariGetDeviceSubtypes",
    "anariGetDeviceExtensions", "anariNewDevice", "anariNewArray1D", "anariNewArray2D", "anariNewArray3D",
    "anariMapArray", "anariUnmapArray", "anariNewLight", "anariNewCamera", "anariNewGeometry", "anariNewSpatialField",
    "anariNewVolume", "anariNewSurface", "anariNewMaterial", "anariNewSampler", "anariNewGroup", "anariNewInstance",
    "anariNewWorld", "anariNewObject", "anariNewRenderer", "anariNewFrame", "anariSetParameter", "anariUnsetParameter",
    "anariUnsetAllParameters", "anariMapParameterArray1D", "anariMapParameterArray2D", "anariMapParameterArray3D",
    "anariUnmapParameterArray", "anariCommitParameters", "anariRelease", "anariRetain", "anariGetObjectSubtypes",
    "anariGetObjectInfo", "anariGetParameterInfo", "anariGetProperty", "anariMapFrame", "anariUnmapFrame",
    "anariRenderFrame", "anariFrameReady", "anariDiscardFrame",
    "makeVisRTXDevice", "visrtxGetObjectExtensions", "visrtxGetInstanceExtensions",
]

StatusCallback = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p)
MemoryDeleter = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
FrameCompletionCallback = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB_PATH)
    for n in API_SYMBOLS:
        getattr(lib, n)
    vp = C.c_void_p
    for n in ("anariLoadLibrary", "anariNewDevice", "anariNewArray1D", "anariNewArray2D", "anariNewArray3D",
              "anariMapArray", "anariNewLight", "anariNewCamera", "anariNewGeometry", "anariNewSpatialField",
              "anariNewVolume", "anariNewSurface", "anariNewMaterial", "anariNewSampler", "anariNewGroup",
              "anariNewInstance", "anariNewWorld", "anariNewObject", "anariNewRenderer", "anariNewFrame",
              "anariMapParameterArray1D", "anariMapParameterArray2D", "anariMapParameterArray3D", "anariMapFrame",
              "makeVisRTXDevice", "anariGetObjectInfo", "anariGetParameterInfo"):
        getattr(lib, n).restype = vp
    lib.anariGetObjectSubtypes.restype = C.POINTER(C.c_char_p)
    lib.anariGetDeviceSubtypes.restype = C.POINTER(C.c_char_p)
    lib.anariGetDeviceExtensions.restype = C.POINTER(C.c_char_p)
    lib.anariGetProperty.restype = C.c_int
    lib.anariFrameReady.restype = C.c_int
    lib.anariLoadLibrary.argtypes = [C.c_char_p, vp, vp]
    lib.anariNewDevice.argtypes = [vp, C.c_char_p]
    lib.makeVisRTXDevice.argtypes = [vp, vp]
    lib.anariNewArray1D.argtypes = [vp, vp, vp, vp, C.c_int, C.c_uint64]
    lib.anariNewArray2D.argtypes = [vp, vp, vp, vp, C.c_int, C.c_uint64, C.c_uint64]
    lib.anariNewArray3D.argtypes = [vp, vp, vp, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
    for n in ("anariNewLight", "anariNewCamera", "anariNewGeometry", "anariNewSpatialField", "anariNewVolume",
              "anariNewMaterial", "anariNewSampler", "anariNewInstance", "anariNewRenderer"):
        getattr(lib, n).argtypes = [vp, C.c_char_p]
    for n in ("anariNewSurface", "anariNewGroup", "anariNewWorld", "anariNewFrame"):
        getattr(lib, n).argtypes = [vp]
    lib.anariSetParameter.argtypes = [vp, vp, C.c_char_p, C.c_int, vp]
    lib.anariUnsetParameter.argtypes = [vp, vp, C.c_char_p]
    lib.anariUnsetAllParameters.argtypes = [vp, vp]
    lib.anariCommitParameters.argtypes = [vp, vp]
    lib.anariRelease.argtypes = [vp, vp]
    lib.anariRetain.argtypes = [vp, vp]
    lib.anariMapArray.argtypes = [vp, vp]
    lib.anariUnmapArray.argtypes = [vp, vp]
    lib.anariGetProperty.argtypes = [vp, vp, C.c_char_p, C.c_int, vp, C.c_uint64, C.c_uint32]
    lib.anariMapFrame.argtypes = [vp, vp, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    lib.anariUnmapFrame.argtypes = [vp, vp, C.c_char_p]
    lib.anariRenderFrame.argtypes = [vp, vp]
    lib.anariFrameReady.argtypes = [vp, vp, C.c_uint32]
    lib.anariDiscardFrame.argtypes = [vp, vp]
    lib.anariGetObjectSubtypes.argtypes = [vp, C.c_int]
    lib.anariGetObjectInfo.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    lib.anariGetParameterInfo.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    lib.anariMapParameterArray1D.argtypes = [vp, vp, C.c_char_p, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.anariUnmapParameterArray.argtypes = [vp, vp, C.c_char_p]
    lib.anariUnloadLibrary.argtypes = [vp]
    return lib


lib = _load()

_NP_OF = {FLOAT32: np.float32, FLOAT64: np.float64, UFIXED8: np.uint8, FIXED8: np.int8, UFIXED16: np.uint16,
          FIXED16: np.int16, FLOAT16: np.float16, UINT8: np.uint8, UINT32: np.uint32}
_PIXEL_NP = {UFIXED8_RGBA_SRGB: (np.uint32, 1), UFIXED8_VEC4: (np.uint32, 1), FLOAT32_VEC4: (np.float32, 4),
             FLOAT32: (np.float32, 1), UINT32: (np.uint32, 1), FLOAT32_VEC3: (np.float32, 3)}


class Device:
    """An ANARIDevice plus the convenience calls of anari_cpp.hpp, as plain methods."""

    def __init__(self, status_cb=None, via_library: bool = True):
        self.messages = []
        self._user_cb = status_cb

        def _cb(user, dev, src, src_type, sev, code, msg):
            m = (sev, code, msg.decode() if msg else "")
            self.messages.append(m)
            if self._user_cb:
                self._user_cb(*m)

        self._cb = StatusCallback(_cb)  # keep alive
        self._keep = []
        if via_library:
            self.library = lib.anariLoadLibrary(b"visrtx_b200", C.cast(self._cb, C.c_void_p), None)
            self.handle = lib.anariNewDevice(self.library, b"default")
        else:
            self.library = None
            self.handle = lib.makeVisRTXDevice(C.cast(self._cb, C.c_void_p), None)
        if not self.handle:
            raise RuntimeError("anariNewDevice failed")

    # ---- objects
    def new(self, kind: str, subtype: Optional[str] = None):
        fn = getattr(lib, "anariNew" + kind)
        return fn(self.handle, subtype.encode()) if subtype is not None else fn(self.handle)

    def new_array1d(self, data: np.ndarray, elem_type: int, n: Optional[int] = None):
        data = np.ascontiguousarray(data)
        self._keep.append(data)
        return lib.anariNewArray1D(self.handle, data.ctypes.data, None, None, elem_type, n or data.shape[0])

    def new_array2d(self, data: np.ndarray, elem_type: int):
        """data: [h, w] or [h, w, channels] (row 0 first)."""
        data = np.ascontiguousarray(data)
        self._keep.append(data)
        return lib.anariNewArray2D(self.handle, data.ctypes.data, None, None, elem_type, data.shape[1], data.shape[0])

    def new_array3d(self, data: np.ndarray, elem_type: int):
        data = np.ascontiguousarray(data)
        self._keep.append(data)
        nz, ny, nx = data.shape
        return lib.anariNewArray3D(self.handle, data.ctypes.data, None, None, elem_type, nx, ny, nz)

    def new_array3d_device(self, dev_ptr: int, elem_type: int, nx: int, ny: int, nz: int):
        """ANARI_NV_ARRAY_CUDA: a shared array over CUDA device memory."""
        return lib.anariNewArray3D(self.handle, dev_ptr, None, None, elem_type, nx, ny, nz)

    def new_object_array(self, handles, elem_type: int):
        arr = (C.c_void_p * len(handles))(*handles)
        self._keep.append(arr)
        return lib.anariNewArray1D(self.handle, C.cast(arr, C.c_void_p), None, None, elem_type, len(handles))

    # ---- parameters
    def set(self, obj, name: str, dtype: int, value):
        if dtype == STRING:
            buf = C.create_string_buffer(value.encode())
        elif dtype in (DATA_TYPE, INT32, BOOL):
            buf = C.c_int32(int(value))
        elif dtype == UINT32:
            buf = C.c_uint32(int(value))
        elif dtype == UINT64:
            buf = C.c_uint64(int(value))
        elif dtype == FLOAT32:
            buf = C.c_float(float(value))
        elif dtype in (VOID_POINTER, FRAME_COMPLETION_CALLBACK, STATUS_CALLBACK) or 500 <= dtype <= 519:
            buf = C.c_void_p(value if isinstance(value, int) or value is None else C.cast(value, C.c_void_p).value)
        elif dtype == UINT32_VEC2:
            buf = (C.c_uint32 * 2)(*[int(v) for v in value])
        elif dtype == FLOAT64_BOX1:
            buf = (C.c_double * 2)(*[float(v) for v in value])
        else:
            vals = [float(v) for v in np.asarray(value, dtype=np.float32).ravel()]
            buf = (C.c_float * len(vals))(*vals)
        lib.anariSetParameter(self.handle, obj, name.encode(), dtype, C.cast(C.pointer(buf), C.c_void_p))

    def unset(self, obj, name: str):
        lib.anariUnsetParameter(self.handle, obj, name.encode())

    def commit(self, obj):
        lib.anariCommitParameters(self.handle, obj)

    def release(self, obj):
        lib.anariRelease(self.handle, obj)

    def retain(self, obj):
        lib.anariRetain(self.handle, obj)

    def map_array(self, arr) -> int:
        return lib.anariMapArray(self.handle, arr)

    def unmap_array(self, arr):
        lib.anariUnmapArray(self.handle, arr)

    # ---- properties
    def get_property(self, obj, name: str, dtype: int, wait: bool = True):
        if dtype == FLOAT32:
            buf = C.c_float()
        elif dtype in (INT32, BOOL):
            buf = C.c_int32()
        elif dtype == FLOAT32_BOX3:
            buf = (C.c_float * 6)()
        elif dtype == FLOAT32_BOX1:
            buf = (C.c_float * 2)()
        elif dtype == STRING_LIST:
            buf = C.POINTER(C.c_char_p)()
        else:
            raise ValueError(dtype)
        ok = lib.anariGetProperty(self.handle, obj, name.encode(), dtype, C.cast(C.pointer(buf), C.c_void_p),
                                  C.sizeof(buf), WAIT if wait else NO_WAIT)
        if not ok:
            return None
        if dtype == STRING_LIST:
            out, i = [], 0
            while buf[i]:
                out.append(buf[i].decode())
                i += 1
            return out
        if dtype in (FLOAT32_BOX3, FLOAT32_BOX1):
            return tuple(buf)
        return buf.value

    def subtypes(self, obj_type: int):
        p = lib.anariGetObjectSubtypes(self.handle, obj_type)
        out, i = [], 0
        while p[i]:
            out.append(p[i].decode())
            i += 1
        return out

    @staticmethod
    def _decode_info(ptr, info_type: int):
        """Turns the pointer an info query returned into a Python value (None when the info does not exist)."""
        if not ptr:
            return None
        if info_type == STRING:
            return C.cast(ptr, C.c_char_p).value.decode()
        if info_type == STRING_LIST:
            p, out, i = C.cast(ptr, C.POINTER(C.c_char_p)), [], 0
            while p[i]:
                out.append(p[i].decode())
                i += 1
            return out
        if info_type == DATA_TYPE_LIST:
            p, out, i = C.cast(ptr, C.POINTER(C.c_int32)), [], 0
            while p[i] != UNKNOWN:
                out.append(p[i])
                i += 1
            return out
        if info_type == PARAMETER_LIST:
            class _P(C.Structure):
                _fields_ = [("name", C.c_char_p), ("type", C.c_int32)]
            p, out, i = C.cast(ptr, C.POINTER(_P)), [], 0
            while p[i].name:
                out.append((p[i].name.decode(), p[i].type))
                i += 1
            return out
        if info_type in (BOOL, INT32, DATA_TYPE):
            return C.cast(ptr, C.POINTER(C.c_int32))[0]
        if info_type == UINT32:
            return C.cast(ptr, C.POINTER(C.c_uint32))[0]
        if info_type == UINT32_VEC2:
            return tuple(C.cast(ptr, C.POINTER(C.c_uint32))[i] for i in range(2))
        n = {FLOAT32: 1, FLOAT32_VEC2: 2, FLOAT32_BOX1: 2, FLOAT32_VEC3: 3, FLOAT32_VEC4: 4, FLOAT32_BOX2: 4,
             FLOAT32_BOX3: 6, FLOAT32_MAT3x4: 12, FLOAT32_MAT4: 16}.get(info_type)
        if n is None:
            raise ValueError(f"no decoder for ANARI type {info_type}")
        v = tuple(C.cast(ptr, C.POINTER(C.c_float))[i] for i in range(n))
        return v[0] if n == 1 else v

    def object_info(self, obj_type: int, subtype, info: str, info_type: int):
        """anariGetObjectInfo: "parameter" (PARAMETER_LIST), "description" / "sourceExtension" (STRING),
        "extension" (STRING_LIST)."""
        st = subtype.encode() if subtype is not None else None
        return self._decode_info(lib.anariGetObjectInfo(self.handle, obj_type, st, info.encode(), info_type), info_type)

    def parameter_info(self, obj_type: int, subtype, name: str, param_type: int, info: str, info_type: int):
        """anariGetParameterInfo: "description", "required", "default", "minimum", "maximum", "value",
        "elementType", "sourceExtension"."""
        st = subtype.encode() if subtype is not None else None
        return self._decode_info(lib.anariGetParameterInfo(self.handle, obj_type, st, name.encode(), param_type,
                                                           info.encode(), info_type), info_type)

    # ---- frames
    def render(self, frame):
        lib.anariRenderFrame(self.handle, frame)

    def wait(self, frame):
        return lib.anariFrameReady(self.handle, frame, WAIT)

    def is_ready(self, frame):
        return bool(lib.anariFrameReady(self.handle, frame, NO_WAIT))

    def map_frame(self, frame, channel: str):
        """returns (numpy copy | device pointer for *CUDA channels, width, height, pixelType)"""
        w, h, t = C.c_uint32(), C.c_uint32(), C.c_int()
        p = lib.anariMapFrame(self.handle, frame, channel.encode(), C.byref(w), C.byref(h), C.byref(t))
        if not p:
            return None, 0, 0, t.value
        if channel.endswith("CUDA") or channel.endswith("GPU"):
            return p, w.value, h.value, t.value
        npdt, comps = _PIXEL_NP[t.value]
        n = w.value * h.value * comps
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(npdt))), shape=(n,)).copy()
        lib.anariUnmapFrame(self.handle, frame, channel.encode())
        return (arr.reshape(-1, comps) if comps > 1 else arr), w.value, h.value, t.value

    def close(self):
        if self.handle:
            lib.anariRelease(self.handle, self.handle)
            self.handle = None
        if self.library:
            lib.anariUnloadLibrary(self.library)
            self.library = None
