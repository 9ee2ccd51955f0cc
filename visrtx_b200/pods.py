"""Plain-old-data mirror of include/dvr_b200.h: enums, ctypes structs and their pure-Python builders.

Nothing here loads a native library or computes anything — the module exists so that code which only needs the
parameter blocks (the oracle bindings under tests/, the reference arm of bench.py) does not pull in libdvr_b200.so.
:mod:`visrtx_b200.capi` re-exports everything below next to the bound entry points.
"""
from __future__ import annotations

import ctypes as C

DVR_TF_SIZE = 256
DVR_MACROCELL_WIDTH = 16

# enums (include/dvr_b200.h)
DVR_OK, DVR_ERR_INVALID_ARGUMENT, DVR_ERR_NO_DEVICE, DVR_ERR_CUDA, DVR_ERR_UNSUPPORTED, DVR_ERR_OUT_OF_MEMORY = (
    0, -1, -2, -3, -4, -5)
DVR_FLOAT32, DVR_UFIXED8, DVR_FIXED8, DVR_UFIXED16, DVR_FIXED16, DVR_FLOAT64, DVR_FLOAT16 = range(7)
DVR_FILTER_LINEAR, DVR_FILTER_NEAREST = 0, 1
DVR_FORMAT_FLOAT32_VEC4, DVR_FORMAT_UFIXED8_VEC4, DVR_FORMAT_UFIXED8_RGBA_SRGB = 0, 1, 2
DVR_CAMERA_PERSPECTIVE, DVR_CAMERA_ORTHOGRAPHIC = 0, 1
DVR_INTEGRATOR_RAYCAST, DVR_INTEGRATOR_DEFAULT, DVR_INTEGRATOR_DPT, DVR_INTEGRATOR_TEST = 0, 1, 2, 3
DVR_SKIP_OFF, DVR_SKIP_ON, DVR_SKIP_AUTO = 0, 1, 2

DVR_IMAGE_FLOAT32, DVR_IMAGE_UFIXED8, DVR_IMAGE_UFIXED16, DVR_IMAGE_UFIXED32, DVR_IMAGE_SRGB8 = range(5)
DVR_GEOMETRY_TRIANGLE, DVR_GEOMETRY_SPHERE = 0, 1
DVR_ALPHA_OPAQUE, DVR_ALPHA_BLEND, DVR_ALPHA_MASK = 0, 1, 2
DVR_LIGHT_DIRECTIONAL, DVR_LIGHT_POINT = 0, 1


IDENTITY_3X4 = (1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0)
_f4 = C.c_float * 4


class DvrCamera(C.Structure):
    _fields_ = [("type", C.c_int32), ("region", C.c_float * 4), ("pos", C.c_float * 3), ("dir", C.c_float * 3),
                ("up", C.c_float * 3), ("du", C.c_float * 3), ("dv", C.c_float * 3), ("p00", C.c_float * 3),
                ("scaledAperture", C.c_float), ("aspect", C.c_float)]


class DvrVolumeInstance(C.Structure):
    _fields_ = [("volume", C.c_void_p), ("worldToObject", C.c_float * 12), ("instanceId", C.c_uint32),
                ("_pad", C.c_uint32)]


class DvrFrameBuffers(C.Structure):
    _fields_ = [("colorAccumulation", C.c_void_p), ("outColor", C.c_void_p), ("depth", C.c_void_p),
                ("primId", C.c_void_p), ("objId", C.c_void_p), ("instId", C.c_void_p), ("albedo", C.c_void_p),
                ("normal", C.c_void_p), ("outColorMirror", C.c_void_p), ("depthMirror", C.c_void_p)]


class DvrFrameParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_int32), ("integrator", C.c_int32),
                ("frameID", C.c_int32), ("checkerboardID", C.c_int32), ("numIterations", C.c_int32),
                ("inverseVolumeSamplingRate", C.c_float), ("background", C.c_float * 4),
                ("tileRank", C.c_uint32), ("tileRanks", C.c_uint32), ("useMacrocellSkipping", C.c_int32),
                ("tileBand", C.c_int32), ("maxDepth", C.c_int32), ("ambientRadiance", C.c_float),
                ("occlusionDistance", C.c_float), ("dptReferenceGrid", C.c_int32), ("partialCullToBounds", C.c_int32), ("_reserved", C.c_int32 * 1),
                ("backgroundImage", C.c_void_p)]


class DvrRenderStats(C.Structure):
    _fields_ = [("samplesTaken", C.c_ulonglong), ("samplesSkipped", C.c_ulonglong), ("raysHit", C.c_ulonglong),
                ("macrocellsTouched", C.c_ulonglong)]


class DvrPeerSync(C.Structure):
    _fields_ = [("nSignal", C.c_uint32), ("signalValue", C.c_uint32), ("signal", C.c_void_p * 16),
                ("nWait", C.c_uint32), ("waitValue", C.c_uint32), ("wait", C.c_void_p), ("errorFlag", C.c_void_p)]


class DvrSlabExchange(C.Structure):
    _fields_ = [("nRanks", C.c_uint32), ("rank", C.c_uint32), ("seq", C.c_uint32), ("maxRegions", C.c_uint32),
                ("partialRgba", C.c_void_p), ("partialDepth", C.c_void_p), ("regionFlags", C.c_void_p),
                ("resolvedFlags", C.c_void_p), ("regionDone", C.c_void_p), ("errorFlag", C.c_void_p),
                ("waitAllResolved", C.c_int32), ("_pad", C.c_int32), ("timing", C.c_void_p)]


class DvrSurfaceDesc(C.Structure):
    _fields_ = [("geometryType", C.c_int32), ("nVertices", C.c_uint32), ("vertexPosition", C.c_void_p),
                ("nPrimitives", C.c_uint32), ("index", C.c_void_p), ("vertexNormal", C.c_void_p),
                ("vertexRadius", C.c_void_p), ("radius", C.c_float), ("primitiveId", C.c_void_p),
                ("cullBackfaces", C.c_int32), ("color", C.c_float * 4), ("opacity", C.c_float),
                ("alphaMode", C.c_int32), ("alphaCutoff", C.c_float), ("surfaceId", C.c_uint32),
                ("instanceId", C.c_uint32), ("objectToWorld", C.c_float * 12)]


class DvrLight(C.Structure):
    _fields_ = [("type", C.c_int32), ("color", C.c_float * 3), ("vec", C.c_float * 3), ("strength", C.c_float)]


class DvrSceneParams(C.Structure):
    _fields_ = [("surfaces", C.c_void_p), ("lights", C.c_void_p), ("nLights", C.c_uint32),
                ("ambientColor", C.c_float * 3), ("ambientRadiance", C.c_float), ("occlusionDistance", C.c_float),
                ("ambientSamples", C.c_int32), ("cullTriangleBackfaces", C.c_int32)]


def surface_descs(surfaces):
    """(DvrSurfaceDesc array, keep-alive list) from dicts with the ANARI parameter names: geometry ('triangle' |
    'sphere'), vertex.position, primitive.index, vertex.normal, vertex.radius, radius, primitive.id, cullBackfaces,
    color (3 or 4), opacity, alphaMode, alphaCutoff, id, instanceId, transform (row-major 3x4 object->world)."""
    import numpy as np
    arr = (DvrSurfaceDesc * max(len(surfaces), 1))()
    keep = []

    def ptr(a, dtype):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype)
        keep.append(a)
        return a.ctypes.data

    for i, sdef in enumerate(surfaces):
        d = arr[i]
        tri = sdef.get("geometry", "triangle") == "triangle"
        d.geometryType = DVR_GEOMETRY_TRIANGLE if tri else DVR_GEOMETRY_SPHERE
        pos = np.ascontiguousarray(sdef["vertex.position"], np.float32).reshape(-1, 3)
        d.nVertices = pos.shape[0]
        d.vertexPosition = ptr(pos, np.float32)
        idx = sdef.get("primitive.index")
        if idx is not None:
            idx = np.ascontiguousarray(idx, np.uint32).reshape(-1, 3 if tri else 1)
            d.nPrimitives = idx.shape[0]
        else:
            d.nPrimitives = pos.shape[0] // 3 if tri else pos.shape[0]
        d.index = ptr(idx, np.uint32)
        d.vertexNormal = ptr(sdef.get("vertex.normal"), np.float32)
        d.vertexRadius = ptr(sdef.get("vertex.radius"), np.float32)
        d.radius = float(sdef.get("radius", 0.01))
        d.primitiveId = ptr(sdef.get("primitive.id"), np.uint32)
        d.cullBackfaces = 1 if sdef.get("cullBackfaces", False) else 0
        col = list(sdef.get("color", (0.8, 0.8, 0.8, 1.0)))
        if len(col) == 3:
            col.append(1.0)
        d.color = _f4(*col)
        d.opacity = float(sdef.get("opacity", 1.0))
        d.alphaMode = {"opaque": DVR_ALPHA_OPAQUE, "blend": DVR_ALPHA_BLEND, "mask": DVR_ALPHA_MASK}[
            sdef.get("alphaMode", "opaque")]
        d.alphaCutoff = float(sdef.get("alphaCutoff", 0.5))
        d.surfaceId = int(sdef.get("id", 0xFFFFFFFF))
        d.instanceId = int(sdef.get("instanceId", 0xFFFFFFFF))
        d.objectToWorld = (C.c_float * 12)(*sdef.get("transform", IDENTITY_3X4))
    return arr, keep


def scene_params(surfaces_handle=None, lights=(), ambient_color=(1.0, 1.0, 1.0), ambient_radiance=0.0,
                 occlusion_distance=1e20, ambient_samples=1, cull_triangle_backfaces=False):
    """(DvrSceneParams, keep-alive) — lights: dicts {type: 'directional'|'point', color, direction|position,
    irradiance|intensity}."""
    p = DvrSceneParams()
    larr = (DvrLight * max(len(lights), 1))()
    for i, l in enumerate(lights):
        point = l.get("type", "directional") == "point"
        larr[i].type = DVR_LIGHT_POINT if point else DVR_LIGHT_DIRECTIONAL
        larr[i].color = (C.c_float * 3)(*l.get("color", (1.0, 1.0, 1.0)))
        larr[i].vec = (C.c_float * 3)(*(l["position"] if point else l["direction"]))
        larr[i].strength = float(l.get("intensity", 1.0) if point else l.get("irradiance", 1.0))
    p.surfaces = surfaces_handle
    p.lights = C.cast(larr, C.c_void_p)
    p.nLights = len(lights)
    p.ambientColor = (C.c_float * 3)(*ambient_color)
    p.ambientRadiance = float(ambient_radiance)
    p.occlusionDistance = float(occlusion_distance)
    p.ambientSamples = int(ambient_samples)
    p.cullTriangleBackfaces = 1 if cull_triangle_backfaces else 0
    return p, larr


def peer_sync(signal_ptrs=(), signal_value=0, wait_ptr=0, n_wait=0, wait_value=0, error_flag=0) -> "DvrPeerSync":
    s = DvrPeerSync()
    s.nSignal, s.signalValue = len(signal_ptrs), int(signal_value) & 0xFFFFFFFF
    for i, p in enumerate(signal_ptrs):
        s.signal[i] = p
    s.nWait, s.waitValue, s.wait = int(n_wait), int(wait_value) & 0xFFFFFFFF, wait_ptr or None
    s.errorFlag = error_flag or None
    return s




def frame_params(width, height, fmt=DVR_FORMAT_UFIXED8_RGBA_SRGB, integrator=DVR_INTEGRATOR_RAYCAST, frame_id=0,
                 checkerboard_id=-1, num_iterations=1, volume_sampling_rate=0.125, background=(0.0, 0.0, 0.0, 1.0),
                 tile_rank=0, tile_ranks=1, skip=False, tile_band=1, max_depth=5, ambient_radiance=1.0,
                 occlusion_distance=1e20, dpt_reference_grid=False, partial_cull_to_bounds=False,
                 background_image=None) -> DvrFrameParams:
    p = DvrFrameParams()
    p.width, p.height, p.format, p.integrator = int(width), int(height), int(fmt), int(integrator)
    p.frameID, p.checkerboardID, p.numIterations = int(frame_id), int(checkerboard_id), int(num_iterations)
    import numpy as np
    p.inverseVolumeSamplingRate = float(np.float32(1.0) / np.float32(volume_sampling_rate))
    p.background = _f4(*background)
    p.tileRank, p.tileRanks = int(tile_rank), int(tile_ranks)
    p.useMacrocellSkipping = DVR_SKIP_AUTO if skip == "auto" else (DVR_SKIP_ON if skip else DVR_SKIP_OFF)
    p.tileBand = int(tile_band)
    p.maxDepth, p.ambientRadiance, p.occlusionDistance = int(max_depth), float(ambient_radiance), float(occlusion_distance)
    p.dptReferenceGrid = 1 if dpt_reference_grid else 0
    p.partialCullToBounds = 1 if partial_cull_to_bounds else 0
    p.backgroundImage = background_image.handle if background_image is not None else None
    return p


def frame_buffers(accum: int, out_color: int, depth: int = 0, prim: int = 0, obj: int = 0, inst: int = 0,
                  albedo: int = 0, normal: int = 0, color_mirror: int = 0, depth_mirror: int = 0) -> DvrFrameBuffers:
    b = DvrFrameBuffers()
    b.outColorMirror = color_mirror or None
    b.depthMirror = depth_mirror or None
    b.colorAccumulation, b.outColor = accum or None, out_color or None
    b.depth, b.primId, b.objId, b.instId = depth or None, prim or None, obj or None, inst or None
    b.albedo, b.normal = albedo or None, normal or None
    return b
