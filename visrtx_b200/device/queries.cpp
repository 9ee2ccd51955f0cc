// queries.cpp — object / parameter introspection of the device (anariGetObjectSubtypes, anariGetObjectInfo,
// anariGetParameterInfo).
//
// The reference answers these calls from code generated out of devices/rtx/visrtx_device.json and the ANARI-SDK's
// extension definitions (VisRTXDeviceQueries.cpp; the generator and the khr_*.json files live in the ANARI-SDK, which
// is not part of the reference tree).  Applications use them to build their parameter editors: TSD walks
// "parameter" -> "description" / "default" / "minimum" / "maximum" / "value" for every renderer subtype
// (tsd/src/tsd/core/Object.cpp:350-424) and reads the "extension" list (AnariObjectCache.cpp:15,
// VisRTXFeatureUtility.cpp:72).  Here the same answers come from hand-written tables that list exactly the
// parameters this device reads (objects.cpp / device.cpp) — renderer entries follow visrtx_device.json:51-311 for
// the parameters the DVR path honours; "default" is always the value the device falls back to when the parameter
// is unset (for the camera that is the reference's own (0,0,1) view direction, camera/Camera.cpp:74).
#include <cstring>
#include <string>
#include <vector>

#include "objects.h"

namespace b200 {
namespace {

struct ParamDesc
{
  const char *name;
  ANARIDataType type;
  const char *description;
  const void *def = nullptr;
  const void *minimum = nullptr;
  const void *maximum = nullptr;
  const char *const *values = nullptr;        // valid strings of an ANARI_STRING parameter
  const ANARIDataType *elementTypes = nullptr; // accepted element types of an array parameter
  bool required = false;
  const char *extension = "ANARI_KHR_CORE";
};

struct ObjectInfo
{
  ANARIDataType type;
  const char *subtype; // nullptr: the object type has no subtypes
  const char *description;
  const char *extension;
  std::vector<ParamDesc> params;
  std::vector<ANARIParameter> list; // name/type pairs + terminator, built once
};

// ---- constants the tables point at -------------------------------------------------------------------------------
const float kZero3[3] = {0.f, 0.f, 0.f}, kOne3[3] = {1.f, 1.f, 1.f}, kDirZ[3] = {0.f, 0.f, 1.f}, kUpY[3] = {0.f, 1.f, 0.f};
const float kRegion[4] = {0.f, 0.f, 1.f, 1.f};
const float kFovy = 60.f * 3.14159265358979323846f / 180.f;
const float kF0 = 0.f, kF1 = 1.f, kF10 = 10.f, kRate = 0.125f, kRateMin = 1e-3f, kFar = 1e20f, kPi = 3.14159265358979323846f;
const float kBackground[4] = {0.f, 0.f, 0.f, 1.f};
const float kRange01[2] = {0.f, 1.f};
const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
const int32_t kI0 = 0, kI1 = 1, kI5 = 5, kI128 = 128, kI256 = 256, kFalse = 0;
const uint32_t kIdNone = ~0u;
const uint32_t kSize10[2] = {10u, 10u};
const char *const kFilters[] = {"linear", "nearest", nullptr};
const char *const kLinear = "linear";

const ANARIDataType kVoxelTypes[] = {ANARI_FLOAT32, ANARI_UFIXED8, ANARI_FIXED8, ANARI_UFIXED16, ANARI_FIXED16,
    ANARI_FLOAT64, ANARI_FLOAT16, ANARI_UNKNOWN};
const ANARIDataType kByteType[] = {ANARI_UINT8, ANARI_UNKNOWN};
const ANARIDataType kColorTypes[] = {ANARI_FLOAT32_VEC3, ANARI_FLOAT32_VEC4, ANARI_UNKNOWN};
const ANARIDataType kFloatType[] = {ANARI_FLOAT32, ANARI_UNKNOWN};
const ANARIDataType kVolumeType[] = {ANARI_VOLUME, ANARI_UNKNOWN};
const ANARIDataType kInstanceType[] = {ANARI_INSTANCE, ANARI_UNKNOWN};
const ANARIDataType kSurfaceType[] = {ANARI_SURFACE, ANARI_UNKNOWN};
const ANARIDataType kLightType[] = {ANARI_LIGHT, ANARI_UNKNOWN};
const ANARIDataType kVec3Type[] = {ANARI_FLOAT32_VEC3, ANARI_UNKNOWN};
const ANARIDataType kUvec3Type[] = {ANARI_UINT32_VEC3, ANARI_UNKNOWN};
const ANARIDataType kUintType[] = {ANARI_UINT32, ANARI_UNKNOWN};
const float kMatteColor[3] = {0.8f, 0.8f, 0.8f}, kDirNegZ[3] = {0.f, 0.f, -1.f}, kRadius = 0.01f, kHalf = 0.5f;
const char *const kAlphaModes[] = {"opaque", "blend", "mask", nullptr};
const char *const kOpaque = "opaque";
const ANARIDataType kColorChannelTypes[] = {ANARI_UFIXED8_VEC4, ANARI_UFIXED8_RGBA_SRGB, ANARI_FLOAT32_VEC4, ANARI_UNKNOWN};

ParamDesc P(const char *name, ANARIDataType type, const char *description, const void *def = nullptr,
    const void *minimum = nullptr, const void *maximum = nullptr)
{
  ParamDesc p;
  p.name = name;
  p.type = type;
  p.description = description;
  p.def = def;
  p.minimum = minimum;
  p.maximum = maximum;
  return p;
}
ParamDesc ext(ParamDesc p, const char *extension)
{
  p.extension = extension;
  return p;
}
ParamDesc req(ParamDesc p)
{
  p.required = true;
  return p;
}
ParamDesc elems(ParamDesc p, const ANARIDataType *types)
{
  p.elementTypes = types;
  return p;
}
ParamDesc strings(ParamDesc p, const char *const *values)
{
  p.values = values;
  return p;
}

const ParamDesc kName = P("name", ANARI_STRING, "optional object name");
const char *const kSortLast = "sortLast";
const char *const kMultiGpuModes[] = {"sortLast", "sortFirst", nullptr};

std::vector<ParamDesc> cameraCommon()
{
  return {kName, P("position", ANARI_FLOAT32_VEC3, "position of the camera in world-space", kZero3),
      P("direction", ANARI_FLOAT32_VEC3, "main viewing direction of the camera", kDirZ),
      P("up", ANARI_FLOAT32_VEC3, "up direction of the camera", kUpY),
      P("imageRegion", ANARI_FLOAT32_BOX2, "region of the sensor in normalized screen-space coordinates", kRegion),
      P("aspect", ANARI_FLOAT32, "ratio of width by height of the frame", &kF1, &kF0)};
}

std::vector<ParamDesc> rendererCommon(bool rate, bool progressive)
{
  std::vector<ParamDesc> v = {kName,
      ext(P("background", ANARI_FLOAT32_VEC4, "background color and alpha (RGBA)", kBackground),
          "ANARI_KHR_RENDERER_BACKGROUND_COLOR")};
  if (progressive) {
    // visrtx_device.json:56-96
    v.push_back(P("sampleLimit", ANARI_INT32, "stop refining the frame after this number of samples", &kI128, &kI0));
    v.push_back(P("checkerboarding", ANARI_BOOL, "use checkerboarding to lower frame latency", &kFalse));
    v.push_back(P("pixelSamples", ANARI_INT32, "samples per-pixel", &kI1, &kI1));
  }
  if (rate) // visrtx_device.json:129-137
    v.push_back(
        P("volumeSamplingRate", ANARI_FLOAT32, "sampling rate of volumes when ray marching", &kRate, &kRateMin, &kF10));
  if (rate) { // surface shading of mixed scenes (visrtx_device.json: ambientColor / ambientRadiance / cullTriangleBackfaces)
    v.push_back(ext(P("ambientColor", ANARI_FLOAT32_VEC3, "ambient light color", kOne3), "ANARI_KHR_RENDERER_AMBIENT_LIGHT"));
    v.push_back(ext(P("ambientRadiance", ANARI_FLOAT32, "ambient light intensity", &kF0, &kF0),
        "ANARI_KHR_RENDERER_AMBIENT_LIGHT"));
    v.push_back(P("cullTriangleBackfaces", ANARI_BOOL, "enable triangle back face culling", &kFalse));
    if (progressive) {
      v.push_back(P("ambientSamples", ANARI_INT32, "number of ambient occlusion samples each frame", &kI1, &kI0, &kI256));
      v.push_back(P("ambientOcclusionDistance", ANARI_FLOAT32, "ambient occlusion distance", &kFar, &kF0));
    }
  }
  // extensions of this device; all image-neutral
  v.push_back(ext(P("macrocellSkipping", ANARI_BOOL,
                      "skip macrocells the transfer function makes fully transparent (unset: decided per volume)"),
      "ANARI_VISRTX_B200_DVR"));
  v.push_back(ext(P("sortFirstRank", ANARI_INT32, "this process' rank when the frame is split by tile rows", &kI0, &kI0),
      "ANARI_VISRTX_B200_DVR"));
  v.push_back(ext(P("sortFirstRanks", ANARI_INT32, "number of processes sharing the frame by tile rows", &kI1, &kI1),
      "ANARI_VISRTX_B200_DVR"));
  return v;
}

std::vector<ObjectInfo> buildTables()
{
  std::vector<ObjectInfo> t;
  // ---- cameras (camera/Camera.cpp:66-80, Perspective.cpp:44-60, Orthographic.cpp:38-42)
  {
    auto p = cameraCommon();
    p.push_back(P("fovy", ANARI_FLOAT32, "vertical field of view in radians", &kFovy, &kF0, &kPi));
    p.push_back(ext(P("focusDistance", ANARI_FLOAT32, "distance at which the image is sharpest", &kF1, &kF0),
        "ANARI_KHR_CAMERA_DEPTH_OF_FIELD"));
    p.push_back(ext(P("apertureRadius", ANARI_FLOAT32, "size of the aperture, controls the depth of field", &kF0, &kF0),
        "ANARI_KHR_CAMERA_DEPTH_OF_FIELD"));
    t.push_back({ANARI_CAMERA, "perspective", "perspective camera", "ANARI_KHR_CAMERA_PERSPECTIVE", p, {}});
  }
  {
    auto p = cameraCommon();
    p.push_back(P("height", ANARI_FLOAT32, "height of the image plane in world units", &kF1, &kF0));
    t.push_back({ANARI_CAMERA, "orthographic", "orthographic camera", "ANARI_KHR_CAMERA_ORTHOGRAPHIC", p, {}});
  }
  // ---- spatial fields (StructuredRegularField.cpp:90-110, NvdbRegularField.cpp:55-75)
  t.push_back({ANARI_SPATIAL_FIELD, "structuredRegular", "structured regular grid of scalars",
      "ANARI_KHR_SPATIAL_FIELD_STRUCTURED_REGULAR",
      {kName, req(elems(P("data", ANARI_ARRAY3D, "array of the vertex-centered scalar values"), kVoxelTypes)),
          P("origin", ANARI_FLOAT32_VEC3, "origin of the grid in object-space", kZero3),
          P("spacing", ANARI_FLOAT32_VEC3, "size of the grid cells in object-space", kOne3),
          strings(P("filter", ANARI_STRING, "filter used for reconstructing the field", &kLinear), kFilters)},
      {}});
  t.push_back({ANARI_SPATIAL_FIELD, "nanovdb", "NanoVDB grid (Float, Fp4, Fp8, Fp16, FpN)",
      "ANARI_VISRTX_SPATIAL_FIELD_NANOVDB",
      {kName, req(elems(P("data", ANARI_ARRAY1D, "one serialized NanoVDB grid"), kByteType))}, {}});
  // ---- volumes (TransferFunction1D.cpp:40-100); "scivis" is the reference's alias of the same object
  for (const char *st : {"transferFunction1D", "scivis"})
    t.push_back({ANARI_VOLUME, st, "scalar field mapped to color and opacity by a 1D transfer function",
        "ANARI_KHR_VOLUME_TRANSFER_FUNCTION1D",
        {kName, req(P("value", ANARI_SPATIAL_FIELD, "spatial field used for the field values of the volume")),
            elems(P("color", ANARI_ARRAY1D, "array to map sampled and clamped field values to color"), kColorTypes),
            elems(P("opacity", ANARI_ARRAY1D, "array to map sampled and clamped field values to opacity"), kFloatType),
            P("valueRange", ANARI_FLOAT32_BOX1, "sampled values are clamped to this range", kRange01),
            P("unitDistance", ANARI_FLOAT32, "distance after which an opacity fraction of the color is absorbed", &kF1,
                &kF0),
            P("id", ANARI_UINT32, "user id written to the objectId channel", &kIdNone)},
        {}});
  // ---- renderers (Renderer.cpp:152-170, Raycast.cpp:45-50, DiffusePathTracer.cpp:41-53)
  for (const char *st : {"default", "ao", "directLight"})
    t.push_back({ANARI_RENDERER, st, "ray-marched volumes, progressive accumulation", "ANARI_KHR_CORE",
        rendererCommon(true, true), {}});
  t.push_back({ANARI_RENDERER, "raycast", "single-shot ray-marched volumes", "ANARI_KHR_CORE",
      rendererCommon(true, false), {}});
  t.push_back({ANARI_RENDERER, "test", "paints the primary ray directions", "ANARI_KHR_CORE", rendererCommon(false, true),
      {}});
  for (const char *st : {"dpt", "diffuse_pathtracer"}) {
    auto p = rendererCommon(false, true);
    p.push_back(P("maxDepth", ANARI_INT32, "maximum number of scattering events per path", &kI5, &kI1, &kI256));
    p.push_back(ext(P("ambientRadiance", ANARI_FLOAT32, "intensity of the ambient light", &kF1, &kF0),
        "ANARI_KHR_RENDERER_AMBIENT_LIGHT"));
    p.push_back(P("ambientOcclusionDistance", ANARI_FLOAT32, "ambient occlusion distance", &kFar, &kF0));
    p.push_back(ext(P("dptReferenceGrid", ANARI_BOOL,
                        "track through the majorant grid as the reference builds it instead of the conservative one",
                        &kFalse),
        "ANARI_VISRTX_B200_DVR"));
    t.push_back({ANARI_RENDERER, st, "volumetric path tracer (delta tracking)", "ANARI_KHR_CORE", p, {}});
  }
  // ---- scene hierarchy (World.cpp, Group.cpp, Instance.cpp)
  t.push_back({ANARI_INSTANCE, "transform", "places a group in the world", "ANARI_KHR_INSTANCE_TRANSFORM",
      {kName, P("group", ANARI_GROUP, "group to be instanced"),
          P("transform", ANARI_FLOAT32_MAT4, "object-to-world transform", kIdentity),
          P("id", ANARI_UINT32, "user id written to the instanceId channel", &kIdNone)},
      {}});
  t.push_back({ANARI_GROUP, nullptr, "container of volumes, surfaces and lights", "ANARI_KHR_CORE",
      {kName, elems(P("volume", ANARI_ARRAY1D, "volumes of the group"), kVolumeType),
          elems(P("surface", ANARI_ARRAY1D, "surfaces of the group"), kSurfaceType),
          elems(P("light", ANARI_ARRAY1D, "lights of the group"), kLightType)},
      {}});
  t.push_back({ANARI_WORLD, nullptr, "container of the scene", "ANARI_KHR_CORE",
      {kName, elems(P("volume", ANARI_ARRAY1D, "volumes placed directly in the world"), kVolumeType),
          elems(P("surface", ANARI_ARRAY1D, "surfaces placed directly in the world"), kSurfaceType),
          elems(P("light", ANARI_ARRAY1D, "lights placed directly in the world"), kLightType),
          elems(P("instance", ANARI_ARRAY1D, "instances of the world"), kInstanceType)},
      {}});
  // ---- surfaces and lights of mixed scenes (Triangle.cpp:44-57, Sphere.cu:47-55, Matte.cpp:38-52, Surface.cpp:42-47,
  //      Directional.cpp:38-47, Point.cpp:38-48)
  t.push_back({ANARI_GEOMETRY, "triangle", "triangle mesh", "ANARI_KHR_GEOMETRY_TRIANGLE",
      {kName, req(elems(P("vertex.position", ANARI_ARRAY1D, "vertex positions"), kVec3Type)),
          elems(P("vertex.normal", ANARI_ARRAY1D, "vertex normals"), kVec3Type),
          elems(P("primitive.index", ANARI_ARRAY1D, "vertex indices of each triangle"), kUvec3Type),
          elems(P("primitive.id", ANARI_ARRAY1D, "user id of each triangle (primitiveId channel)"), kUintType),
          P("cullBackfaces", ANARI_BOOL, "cull back-facing triangles on primary rays", &kFalse)},
      {}});
  t.push_back({ANARI_GEOMETRY, "sphere", "spheres", "ANARI_KHR_GEOMETRY_SPHERE",
      {kName, req(elems(P("vertex.position", ANARI_ARRAY1D, "sphere centers"), kVec3Type)),
          elems(P("vertex.radius", ANARI_ARRAY1D, "per-sphere radius"), kFloatType),
          elems(P("primitive.index", ANARI_ARRAY1D, "index of each sphere's center"), kUintType),
          elems(P("primitive.id", ANARI_ARRAY1D, "user id of each sphere (primitiveId channel)"), kUintType),
          P("radius", ANARI_FLOAT32, "radius of spheres without a vertex.radius", &kRadius, &kF0)},
      {}});
  t.push_back({ANARI_MATERIAL, "matte", "diffuse material (constant color and opacity)", "ANARI_KHR_MATERIAL_MATTE",
      {kName, P("color", ANARI_FLOAT32_VEC3, "diffuse color", kMatteColor),
          P("opacity", ANARI_FLOAT32, "opacity", &kF1, &kF0, &kF1),
          strings(P("alphaMode", ANARI_STRING, "how opacity is interpreted", &kOpaque), kAlphaModes),
          P("alphaCutoff", ANARI_FLOAT32, "threshold of alphaMode mask", &kHalf, &kF0, &kF1)},
      {}});
  t.push_back({ANARI_SURFACE, nullptr, "geometry with a material", "ANARI_KHR_CORE",
      {kName, req(P("geometry", ANARI_GEOMETRY, "geometry of the surface")),
          req(P("material", ANARI_MATERIAL, "material of the surface")),
          P("id", ANARI_UINT32, "user id written to the objectId channel", &kIdNone)},
      {}});
  t.push_back({ANARI_LIGHT, "directional", "distant light", "ANARI_KHR_LIGHT_DIRECTIONAL",
      {kName, P("color", ANARI_FLOAT32_VEC3, "color of the light", kOne3),
          P("direction", ANARI_FLOAT32_VEC3, "direction the light travels in", kDirNegZ),
          P("irradiance", ANARI_FLOAT32, "irradiance on a surface facing the light", &kF1, &kF0)},
      {}});
  t.push_back({ANARI_LIGHT, "point", "point light", "ANARI_KHR_LIGHT_POINT",
      {kName, P("color", ANARI_FLOAT32_VEC3, "color of the light", kOne3),
          P("position", ANARI_FLOAT32_VEC3, "position of the light", kZero3),
          P("intensity", ANARI_FLOAT32, "radiant intensity", &kF1, &kF0)},
      {}});
  // ---- frame (Frame.cu:120-200)
  t.push_back({ANARI_FRAME, nullptr, "render target", "ANARI_KHR_CORE",
      {kName, req(P("size", ANARI_UINT32_VEC2, "size of the frame in pixels", kSize10)),
          req(elems(P("channel.color", ANARI_DATA_TYPE, "type of the color channel"), kColorChannelTypes)),
          P("channel.depth", ANARI_DATA_TYPE, "enables the depth channel (ANARI_FLOAT32)"),
          ext(P("channel.primitiveId", ANARI_DATA_TYPE, "enables the primitive id channel (ANARI_UINT32)"),
              "ANARI_KHR_FRAME_CHANNEL_PRIMITIVE_ID"),
          ext(P("channel.objectId", ANARI_DATA_TYPE, "enables the object id channel (ANARI_UINT32)"),
              "ANARI_KHR_FRAME_CHANNEL_OBJECT_ID"),
          ext(P("channel.instanceId", ANARI_DATA_TYPE, "enables the instance id channel (ANARI_UINT32)"),
              "ANARI_KHR_FRAME_CHANNEL_INSTANCE_ID"),
          ext(P("channel.albedo", ANARI_DATA_TYPE, "enables the albedo channel (ANARI_FLOAT32_VEC3)"),
              "ANARI_KHR_FRAME_CHANNEL_ALBEDO"),
          ext(P("channel.normal", ANARI_DATA_TYPE, "enables the normal channel (ANARI_FLOAT32_VEC3)"),
              "ANARI_KHR_FRAME_CHANNEL_NORMAL"),
          req(P("renderer", ANARI_RENDERER, "renderer used to render the frame")),
          req(P("camera", ANARI_CAMERA, "camera used to render the frame")),
          req(P("world", ANARI_WORLD, "world to be rendered")),
          ext(P("frameCompletionCallback", ANARI_FRAME_COMPLETION_CALLBACK, "called when the frame has completed"),
              "ANARI_KHR_FRAME_COMPLETION_CALLBACK"),
          ext(P("frameCompletionCallbackUserData", ANARI_VOID_POINTER, "passed to the completion callback"),
              "ANARI_KHR_FRAME_COMPLETION_CALLBACK")},
      {}});
  // ---- the device itself (VisRTXDevice.cpp:455-470)
  t.push_back({ANARI_DEVICE, nullptr, "B200 direct-volume-rendering device", "ANARI_KHR_CORE",
      {kName, P("statusCallback", ANARI_STATUS_CALLBACK, "callback used to report information to the application"),
          P("statusCallbackUserData", ANARI_VOID_POINTER, "passed to the status callback"),
          P("cudaDevice", ANARI_INT32, "ordinal of the CUDA device to render on", &kI0, &kI0),
          P("forceInit", ANARI_BOOL, "initialise CUDA when the device is committed instead of at first use", &kFalse),
          ext(P("cudaDevices", ANARI_STRING,
                  "multi-GPU: comma separated CUDA device ordinals, display GPU first (also accepted as an Array1D of INT32)"),
              "ANARI_VISRTX_B200_DVR"),
          ext(strings(P("multiGpuMode", ANARI_STRING,
                          "multi-GPU: sortLast = z-slabs of structuredRegular fields, fused march + exchange per GPU; "
                          "sortFirst = fields replicated, interleaved tile rows",
                          &kSortLast),
                  kMultiGpuModes),
              "ANARI_VISRTX_B200_DVR")},
      {}});
  for (ObjectInfo &o : t) {
    for (const ParamDesc &p : o.params)
      o.list.push_back({p.name, p.type});
    o.list.push_back({nullptr, ANARI_UNKNOWN});
  }
  return t;
}

const std::vector<ObjectInfo> &tables()
{
  static const std::vector<ObjectInfo> t = buildTables();
  return t;
}

const ObjectInfo *findObject(ANARIDataType type, const char *subtype)
{
  const ObjectInfo *firstOfType = nullptr;
  for (const ObjectInfo &o : tables()) {
    if (o.type != type)
      continue;
    if (!o.subtype)
      return &o; // no subtypes: whatever the caller passes
    if (subtype && std::strcmp(o.subtype, subtype) == 0)
      return &o;
    if (!firstOfType)
      firstOfType = &o;
  }
  // "default" / NULL on a type whose subtypes are all named: the first one (renderers do have a "default")
  if (firstOfType && (!subtype || !*subtype || std::strcmp(subtype, "default") == 0))
    return firstOfType;
  return nullptr;
}

} // namespace

const char **querySubtypes(ANARIDataType type)
{
  // one NULL-terminated list per object type, in table order; aliases the reference also accepts but does not
  // advertise (diffuse_pathtracer) are left out
  static std::vector<std::pair<ANARIDataType, std::vector<const char *>>> lists = [] {
    std::vector<std::pair<ANARIDataType, std::vector<const char *>>> l;
    for (const ObjectInfo &o : tables()) {
      if (!o.subtype || std::strcmp(o.subtype, "diffuse_pathtracer") == 0)
        continue;
      auto it = l.begin();
      for (; it != l.end() && it->first != o.type; ++it)
        ;
      if (it == l.end()) {
        l.push_back({o.type, {}});
        it = l.end() - 1;
      }
      it->second.push_back(o.subtype);
    }
    for (auto &e : l)
      e.second.push_back(nullptr);
    return l;
  }();
  static const char *none[] = {nullptr};
  for (auto &e : lists)
    if (e.first == type)
      return e.second.data();
  return none;
}

const void *queryObjectInfo(ANARIDataType type, const char *subtype, const char *infoName, ANARIDataType infoType,
    const char **extensions)
{
  if (!infoName)
    return nullptr;
  const std::string n = infoName;
  if (n == "extension" && infoType == ANARI_STRING_LIST)
    return extensions;
  const ObjectInfo *o = findObject(type, subtype);
  if (!o)
    return nullptr;
  if (n == "parameter" && infoType == ANARI_PARAMETER_LIST)
    return o->list.data();
  if (n == "description" && infoType == ANARI_STRING)
    return o->description;
  if (n == "sourceExtension" && infoType == ANARI_STRING)
    return o->extension;
  return nullptr;
}

const void *queryParameterInfo(ANARIDataType type, const char *subtype, const char *paramName, ANARIDataType paramType,
    const char *infoName, ANARIDataType infoType)
{
  if (!paramName || !infoName)
    return nullptr;
  const ObjectInfo *o = findObject(type, subtype);
  if (!o)
    return nullptr;
  const ParamDesc *p = nullptr;
  for (const ParamDesc &c : o->params)
    if (c.type == paramType && std::strcmp(c.name, paramName) == 0) {
      p = &c;
      break;
    }
  if (!p)
    return nullptr;
  static const int32_t yes = 1, no = 0;
  const std::string n = infoName;
  if (n == "description" && infoType == ANARI_STRING)
    return p->description;
  if (n == "required" && infoType == ANARI_BOOL)
    return p->required ? &yes : &no;
  if (n == "sourceExtension" && infoType == ANARI_STRING)
    return p->extension;
  if (n == "value" && infoType == ANARI_STRING_LIST)
    return p->values;
  if (n == "elementType" && infoType == ANARI_DATA_TYPE_LIST)
    return p->elementTypes;
  // typed infos: the caller must ask for them in the parameter's own type.  An ANARI_STRING default is the
  // C string itself (the table stores a pointer to it).
  if (infoType == p->type) {
    const void *v = n == "default" ? p->def : n == "minimum" ? p->minimum : n == "maximum" ? p->maximum : nullptr;
    if (v && p->type == ANARI_STRING)
      return *static_cast<const char *const *>(v);
    return v;
  }
  return nullptr;
}

} // namespace b200
