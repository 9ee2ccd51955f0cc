"""CPU suite, part 3: the ANARI device library — symbols, object/parameter/lifetime semantics that need
no GPU, introspection tables, error reporting through the status callback (never exceptions), and the
reference's multi-threaded object-creation smoke test (tests/api/TestMultiThreadedObjectCreation.cpp)."""
import ctypes as C
import os
import re
import threading

import numpy as np
import pytest

from conftest import HAS_GPU
from visrtx_b200 import anari as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_entry_point():
    hdr = open(os.path.join(ROOT, "include", "anari", "anari.h")).read()
    hdr += open(os.path.join(ROOT, "include", "anari", "ext", "visrtx_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(anari[A-Z][A-Za-z0-9]+|makeVisRTXDevice|visrtxGet[A-Za-z]+)\s*\(", hdr))
    assert declared == set(A.API_SYMBOLS)
    for n in declared:
        assert hasattr(A.lib, n)


def test_load_library_names_and_device_subtypes():
    d = A.Device()
    assert A.lib.anariLoadLibrary(b"no_such_library", None, None) is None
    p = A.lib.anariGetDeviceSubtypes(d.library)
    assert p[0] == b"default" and not p[1]
    ext = d.get_property(d.handle, "extension", A.STRING_LIST)
    for e in ("ANARI_KHR_SPATIAL_FIELD_STRUCTURED_REGULAR", "ANARI_KHR_VOLUME_TRANSFER_FUNCTION1D",
              "ANARI_KHR_FRAME_ACCUMULATION", "ANARI_NV_FRAME_BUFFERS_CUDA", "ANARI_NV_ARRAY_CUDA"):
        assert e in ext
    assert d.get_property(d.handle, "version.major", A.INT32) == 0
    d.close()


def test_direct_device_construction():
    d = A.Device(via_library=False)  # makeVisRTXDevice, visrtx.h:46-52
    assert d.handle
    d.close()


def test_object_subtypes_introspection():
    d = A.Device()
    assert d.subtypes(A.CAMERA) == ["perspective", "orthographic"]
    assert d.subtypes(A.SPATIAL_FIELD) == ["structuredRegular", "nanovdb"]
    assert "transferFunction1D" in d.subtypes(A.VOLUME) and "scivis" in d.subtypes(A.VOLUME)
    assert {"default", "raycast"} <= set(d.subtypes(A.RENDERER))
    assert d.subtypes(A.WORLD) == []
    d.close()


def test_parameters_and_lifetime_without_gpu():
    d = A.Device()
    cam = d.new("Camera", "perspective")
    d.set(cam, "position", A.FLOAT32_VEC3, (1, 2, 3))
    d.set(cam, "fovy", A.FLOAT32, 0.5)
    d.unset(cam, "fovy")
    d.commit(cam)
    d.retain(cam)
    d.release(cam)
    d.release(cam)
    vol = d.new("Volume", "transferFunction1D")
    col = d.new_array1d(np.array([[1, 0, 0], [0, 0, 1]], np.float32), A.FLOAT32_VEC3)
    d.set(vol, "color", A.ARRAY1D, col)
    d.release(col)  # the volume's parameter keeps it alive (INTERNAL ref) and privatises the shared data
    d.commit(vol)
    d.release(vol)
    assert not [m for m in d.messages if m[0] <= A.SEVERITY_ERROR]
    d.close()


def test_managed_array_is_zero_initialised_and_mappable():
    d = A.Device()
    a = A.lib.anariNewArray1D(d.handle, None, None, None, A.FLOAT32, 16)
    p = d.map_array(a)
    v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(16,))
    assert np.all(v == 0)
    v[:] = 3.0
    d.unmap_array(a)
    p2 = d.map_array(a)
    assert np.all(np.ctypeslib.as_array(C.cast(p2, C.POINTER(C.c_float)), shape=(16,)) == 3.0)
    d.unmap_array(a)
    d.release(a)
    d.close()


def test_mapping_an_object_array_keeps_its_elements_alive():
    """The normal ANARI pattern: create instances, put them in an array, release the handles — the array is then the
    only owner.  Mapping that array must not drop those references (the elements would be deleted while World still
    reads them); unmap reconciles the old handle list with the new contents."""
    d = A.Device()
    groups = [d.new("Group") for _ in range(3)]
    insts = []
    for g in groups:
        i = d.new("Instance", "transform")
        d.set(i, "group", A.GROUP, g)
        d.commit(i)
        insts.append(i)
    arr = d.new_object_array(insts[:2], A.INSTANCE)
    for i in insts[:2]:
        d.release(i)  # the array holds the only reference now
    world = d.new("World")
    d.set(world, "instance", A.ARRAY1D, arr)
    d.commit(world)
    for _ in range(3):  # repeated map/unmap without edits must be neutral
        p = d.map_array(arr)
        assert p
        d.unmap_array(arr)
    p = d.map_array(arr)
    handles = C.cast(p, C.POINTER(C.c_void_p))
    assert handles[0] == insts[0] and handles[1] == insts[1]
    d.set(insts[0], "id", A.UINT32, 7)  # still a live object: parameter stores on a freed object would corrupt the heap
    d.commit(insts[0])
    handles[1] = insts[2]  # replace one element while mapped
    d.unmap_array(arr)
    d.release(insts[2])
    d.commit(world)
    lo_hi = d.get_property(world, "bounds", A.FLOAT32_BOX3)  # walks the instance array (flatten)
    assert lo_hi is not None
    for o in [world, arr] + groups:
        d.release(o)
    assert not [m for m in d.messages if m[0] <= A.SEVERITY_ERROR]
    d.close()


def test_captured_array_deleter_runs_on_release():
    d = A.Device()
    calls = []
    data = np.arange(8, dtype=np.float32)
    deleter = A.MemoryDeleter(lambda user, mem: calls.append(mem))
    a = A.lib.anariNewArray1D(d.handle, data.ctypes.data, C.cast(deleter, C.c_void_p), None, A.FLOAT32, 8)
    d.release(a)
    assert calls == [data.ctypes.data]
    d.close()


def test_errors_are_reported_not_thrown():
    d = A.Device()
    bad = A.lib.anariNewArray1D(d.handle, None, None, None, 424242, 4)
    assert bad is None
    assert any(m[0] == A.SEVERITY_ERROR for m in d.messages)
    d.messages.clear()
    # frame with nothing attached: render is skipped with an error message
    f = d.new("Frame")
    d.commit(f)
    d.render(f)
    assert any("skipping render" in m[2] or "failed" in m[2] or "CUDA" in m[2] for m in d.messages)
    assert d.wait(f) == 1
    d.release(f)
    d.close()


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device path")
def test_device_init_failure_is_sticky_and_loud():
    d = A.Device()
    d.set(d.handle, "forceInit", A.BOOL, 1)
    d.commit(d.handle)
    assert any(m[0] == A.SEVERITY_FATAL_ERROR and "no CPU fallback" in m[2] for m in d.messages)
    n = len(d.messages)
    f = d.new("Frame")
    r = d.new("Renderer", "default")
    c = d.new("Camera", "perspective")
    w = d.new("World")
    for k, t, o in (("renderer", A.RENDERER, r), ("camera", A.CAMERA, c), ("world", A.WORLD, w)):
        d.set(f, k, t, o)
    d.commit(f)
    d.render(f)
    assert len(d.messages) > n and any("failed to init" in m[2] for m in d.messages[n:])
    d.close()


def test_multithreaded_object_creation():
    """4 threads x 100 x {camera, world, frame, renderer, volume} create/release concurrently."""
    d = A.Device()
    errs = []

    def work():
        try:
            for _ in range(100):
                objs = [d.new("Camera", "perspective"), d.new("World"), d.new("Frame"), d.new("Renderer", "default"),
                        d.new("Volume", "transferFunction1D"), d.new("Light", "directional"),
                        d.new("Geometry", "triangle")]
                for o in objs:
                    d.commit(o)
                for o in objs:
                    d.release(o)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert not [m for m in d.messages if m[0] == A.SEVERITY_FATAL_ERROR]
    d.close()


def test_parameter_introspection_as_tsd_walks_it():
    """tsd/src/tsd/core/Object.cpp:350-424 (parseANARIObjectInfo): for every renderer subtype the application asks for
    the parameter list and then, per parameter, for description / default / minimum / maximum / value.  Values for the
    renderers are those of devices/rtx/visrtx_device.json; every listed parameter must be answerable."""
    d = A.Device()
    for st in d.subtypes(A.RENDERER):
        params = d.object_info(A.RENDERER, st, "parameter", A.PARAMETER_LIST)
        assert params and ("background", A.FLOAT32_VEC4) in params, st
        for name, t in params:
            desc = d.parameter_info(A.RENDERER, st, name, t, "description", A.STRING)
            assert isinstance(desc, str) and desc, (st, name)
            if t != A.STRING:
                d.parameter_info(A.RENDERER, st, name, t, "default", t)  # decodable or None
    info = lambda st, n, t, what: d.parameter_info(A.RENDERER, st, n, t, what, t)
    # visrtx_device.json:56-64, 88-96, 129-137
    assert (info("default", "sampleLimit", A.INT32, "default"), info("default", "sampleLimit", A.INT32, "minimum")) == (128, 0)
    assert info("default", "sampleLimit", A.INT32, "maximum") is None
    assert (info("default", "pixelSamples", A.INT32, "default"), info("default", "pixelSamples", A.INT32, "minimum")) == (1, 1)
    rate = [info("default", "volumeSamplingRate", A.FLOAT32, k) for k in ("default", "minimum", "maximum")]
    assert rate == [0.125, np.float32(0.001), 10.0]
    assert info("raycast", "volumeSamplingRate", A.FLOAT32, "default") == 0.125
    assert info("default", "checkerboarding", A.BOOL, "default") == 0
    # the single-shot renderer has no accumulation parameters, the path tracer no marching rate
    names = lambda st: [n for n, _ in d.object_info(A.RENDERER, st, "parameter", A.PARAMETER_LIST)]
    assert "sampleLimit" not in names("raycast") and "pixelSamples" not in names("raycast")
    assert "volumeSamplingRate" not in names("dpt") and "maxDepth" in names("dpt")
    assert info("dpt", "ambientRadiance", A.FLOAT32, "default") == 1.0  # DiffusePathTracer.cpp:41
    assert info("dpt", "maxDepth", A.INT32, "default") == 5
    # asking in another type than the parameter's, for an unknown parameter or an unknown subtype: no info
    assert d.parameter_info(A.RENDERER, "default", "sampleLimit", A.INT32, "default", A.FLOAT32) is None
    assert d.parameter_info(A.RENDERER, "default", "sampleLimit", A.FLOAT32, "default", A.FLOAT32) is None
    assert d.parameter_info(A.RENDERER, "default", "noSuchParameter", A.INT32, "default", A.INT32) is None
    assert d.object_info(A.RENDERER, "debug", "parameter", A.PARAMETER_LIST) is None
    # NULL subtype on a renderer means its default renderer
    assert d.object_info(A.RENDERER, None, "parameter", A.PARAMETER_LIST) == d.object_info(
        A.RENDERER, "default", "parameter", A.PARAMETER_LIST)
    d.close()


def test_introspection_of_the_scene_objects_matches_what_the_device_reads():
    d = A.Device()
    par = lambda t, st: dict(d.object_info(t, st, "parameter", A.PARAMETER_LIST))
    cam = par(A.CAMERA, "perspective")
    assert cam["fovy"] == A.FLOAT32 and cam["imageRegion"] == A.FLOAT32_BOX2 and "height" not in cam
    assert "height" in par(A.CAMERA, "orthographic") and "fovy" not in par(A.CAMERA, "orthographic")
    pi = lambda t, st, n, pt, what, it=None: d.parameter_info(t, st, n, pt, what, pt if it is None else it)
    assert pi(A.CAMERA, "perspective", "fovy", A.FLOAT32, "default") == np.float32(np.pi / 3)
    assert pi(A.CAMERA, "perspective", "direction", A.FLOAT32_VEC3, "default") == (0.0, 0.0, 1.0)  # Camera.cpp:74
    assert pi(A.CAMERA, "perspective", "imageRegion", A.FLOAT32_BOX2, "default") == (0.0, 0.0, 1.0, 1.0)
    # fields: required data array and its element types, filter values
    assert pi(A.SPATIAL_FIELD, "structuredRegular", "data", A.ARRAY3D, "required", A.BOOL) == 1
    et = pi(A.SPATIAL_FIELD, "structuredRegular", "data", A.ARRAY3D, "elementType", A.DATA_TYPE_LIST)
    assert {A.FLOAT32, A.UFIXED8, A.UFIXED16, A.FIXED16, A.FLOAT64} <= set(et)
    assert pi(A.SPATIAL_FIELD, "structuredRegular", "filter", A.STRING, "value", A.STRING_LIST) == ["linear", "nearest"]
    assert pi(A.SPATIAL_FIELD, "structuredRegular", "filter", A.STRING, "default") == "linear"
    assert pi(A.SPATIAL_FIELD, "structuredRegular", "spacing", A.FLOAT32_VEC3, "default") == (1.0, 1.0, 1.0)
    assert pi(A.SPATIAL_FIELD, "nanovdb", "data", A.ARRAY1D, "elementType", A.DATA_TYPE_LIST) == [A.UINT8]
    assert pi(A.SPATIAL_FIELD, "structuredRegular", "origin", A.FLOAT32_VEC3, "required", A.BOOL) == 0
    # volume: both names of the transfer-function volume answer alike
    for st in ("transferFunction1D", "scivis"):
        v = par(A.VOLUME, st)
        assert v["value"] == A.SPATIAL_FIELD and v["valueRange"] == A.FLOAT32_BOX1 and v["unitDistance"] == A.FLOAT32
        assert pi(A.VOLUME, st, "valueRange", A.FLOAT32_BOX1, "default") == (0.0, 1.0)
        assert pi(A.VOLUME, st, "color", A.ARRAY1D, "elementType", A.DATA_TYPE_LIST) == [A.FLOAT32_VEC3, A.FLOAT32_VEC4]
    # objects without subtypes answer for any subtype string
    assert par(A.FRAME, None)["channel.color"] == A.DATA_TYPE and par(A.FRAME, "")["size"] == A.UINT32_VEC2
    assert pi(A.FRAME, None, "size", A.UINT32_VEC2, "default") == (10, 10)
    assert set(par(A.WORLD, None)) == {"name", "volume", "surface", "light", "instance"}
    # mixed scenes: the geometry / material / light subtypes the device renders
    assert d.subtypes(A.GEOMETRY) == ["triangle", "sphere"] and d.subtypes(A.MATERIAL) == ["matte"]
    assert d.subtypes(A.LIGHT) == ["directional", "point"]
    assert par(A.GEOMETRY, "triangle")["primitive.index"] == A.ARRAY1D
    assert pi(A.GEOMETRY, "triangle", "primitive.index", A.ARRAY1D, "elementType", A.DATA_TYPE_LIST) == [A.UINT32_VEC3]
    assert abs(pi(A.GEOMETRY, "sphere", "radius", A.FLOAT32, "default") - 0.01) < 1e-9
    assert pi(A.MATERIAL, "matte", "alphaMode", A.STRING, "value", A.STRING_LIST) == ["opaque", "blend", "mask"]
    assert pi(A.LIGHT, "directional", "direction", A.FLOAT32_VEC3, "default") == (0.0, 0.0, -1.0)
    assert pi(A.RENDERER, "default", "ambientSamples", A.INT32, "default") == 1
    assert set(par(A.SURFACE, None)) == {"name", "geometry", "material", "id"}
    assert pi(A.INSTANCE, "transform", "transform", A.FLOAT32_MAT4, "default")[::5] == (1.0, 1.0, 1.0, 1.0)
    # descriptions and source extensions
    assert "NanoVDB" in d.object_info(A.SPATIAL_FIELD, "nanovdb", "description", A.STRING)
    assert d.object_info(A.CAMERA, "orthographic", "sourceExtension", A.STRING) == "ANARI_KHR_CAMERA_ORTHOGRAPHIC"
    assert pi(A.CAMERA, "perspective", "apertureRadius", A.FLOAT32, "sourceExtension",
              A.STRING) == "ANARI_KHR_CAMERA_DEPTH_OF_FIELD"
    exts = d.object_info(A.RENDERER, "default", "extension", A.STRING_LIST)
    assert "ANARI_KHR_SPATIAL_FIELD_STRUCTURED_REGULAR" in exts and "ANARI_NV_ARRAY_CUDA" in exts
    # the device's own parameters (VisRTXDevice.cpp:455-470)
    dev = par(A.DEVICE, "default")
    assert dev["cudaDevice"] == A.INT32 and dev["forceInit"] == A.BOOL and dev["statusCallback"] == A.STATUS_CALLBACK
    assert pi(A.DEVICE, None, "cudaDevice", A.INT32, "default") == 0
    # every advertised subtype of every object type has a parameter list
    for t in (A.CAMERA, A.SPATIAL_FIELD, A.VOLUME, A.RENDERER, A.INSTANCE, A.GEOMETRY, A.MATERIAL, A.LIGHT):
        for st in d.subtypes(t):
            assert d.object_info(t, st, "parameter", A.PARAMETER_LIST), (t, st)
    d.close()


def test_surface_objects_validate_their_parameters_and_extend_the_world_bounds():
    """Geometry / Material / Surface / Light objects of mixed scenes: commit-time validation with the
    reference's messages (Triangle.cpp:59-78, Surface.cpp:49-58) and World 'bounds' over surfaces — no GPU needed."""
    d = A.Device()
    g = d.new("Geometry", "triangle")
    d.set(g, "vertex.position", A.ARRAY1D, d.new_array1d(np.zeros((4, 3), np.float32), A.FLOAT32_VEC3))
    d.commit(g)
    s = d.new("Surface")
    d.set(s, "geometry", A.GEOMETRY, g)
    d.commit(s)
    w = d.new("World")
    d.set(w, "surface", A.ARRAY1D, d.new_object_array([s], A.SURFACE))
    d.commit(w)
    b = d.get_property(w, "bounds", A.FLOAT32_BOX3)
    msgs = " | ".join(m[2] for m in d.messages)
    assert "non-multiple of 3" in msgs and "missing 'material' on ANARISurface" in msgs
    assert b[0] > b[3]  # nothing valid in the world: the empty box
    # a valid sphere surface under an instance transform
    g2 = d.new("Geometry", "sphere")
    d.set(g2, "vertex.position", A.ARRAY1D, d.new_array1d(np.array([[1, 2, 3], [-1, 0, 0]], np.float32), A.FLOAT32_VEC3))
    d.set(g2, "radius", A.FLOAT32, 0.5)
    d.commit(g2)
    m = d.new("Material", "matte")
    d.commit(m)
    s2 = d.new("Surface")
    d.set(s2, "geometry", A.GEOMETRY, g2)
    d.set(s2, "material", A.MATERIAL, m)
    d.commit(s2)
    grp = d.new("Group")
    d.set(grp, "surface", A.ARRAY1D, d.new_object_array([s2], A.SURFACE))
    d.commit(grp)
    inst = d.new("Instance", "transform")
    d.set(inst, "group", A.GROUP, grp)
    d.set(inst, "transform", A.FLOAT32_MAT3x4, (1, 0, 0, 0, 1, 0, 0, 0, 1, 10, 0, 0))  # translate x by 10
    d.commit(inst)
    d.set(w, "instance", A.ARRAY1D, d.new_object_array([inst], A.INSTANCE))
    d.commit(w)
    b = d.get_property(w, "bounds", A.FLOAT32_BOX3)
    assert np.allclose(b, (8.5, -0.5, -0.5, 11.5, 2.5, 3.5))
    d2 = d.new("Geometry", "cone")
    d.commit(d2)
    lt = d.new("Light", "hdri")
    d.commit(lt)
    d.get_property(w, "bounds", A.FLOAT32_BOX3)
    msgs = " | ".join(m[2] for m in d.messages)
    assert "geometry subtype 'cone' is not rendered" in msgs and "light subtype 'hdri' is not rendered" in msgs
    d.close()
