// objects.h — host object model of the B200 DVR ANARI device.
//
// Mirrors the part of VisRTX's helium-based object model that the DVR path touches
// (devices/rtx/{Object,RegisteredObject}.h, array/, camera/, scene/volume/**, scene/{World,Group,Instance}.*,
// renderer/Renderer.*, frame/Frame.*): string-keyed parameter store, deferred commit
// (commitParameters + finalize at the next renderFrame / WAIT query), PUBLIC/INTERNAL reference counts,
// array change observers, accumulation reset on any finalisation.  All GPU work goes through the
// extern "C" launch layer of include/dvr_b200.h.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include <anari/anari.h>
#include "../../include/dvr_b200.h"

namespace b200 {

struct Device;
struct Object;

size_t sizeOfType(ANARIDataType t);
bool isObjectType(ANARIDataType t);
const char *typeName(ANARIDataType t);

struct Param
{
  ANARIDataType type = ANARI_UNKNOWN;
  std::vector<uint8_t> bytes; // POD value (or the string incl. terminator)
  Object *object = nullptr;   // object-typed parameter (holds an INTERNAL ref)
};

// helium::RefType
enum class RefType
{
  PUBLIC,
  INTERNAL
};

struct Object
{
  Object(Device *d, ANARIDataType type, std::string subtype = "");
  virtual ~Object();

  // lifetime (helium::RefCounted): destroyed when both counts reach zero
  void refInc(RefType t);
  void refDec(RefType t);
  int useCount(RefType t) const { return t == RefType::PUBLIC ? m_public.load() : m_internal.load(); }

  // parameters
  void setParam(const char *name, ANARIDataType type, const void *mem);
  void unsetParam(const char *name);
  void unsetAllParams();
  const Param *findParam(const std::string &name) const;
  template <typename T>
  T getParam(const std::string &name, ANARIDataType type, T def) const
  {
    const Param *p = findParam(name);
    if (!p || p->type != type || p->bytes.size() < sizeof(T))
      return def;
    T v;
    std::memcpy(&v, p->bytes.data(), sizeof(T));
    return v;
  }
  bool getParamRaw(const std::string &name, ANARIDataType type, void *dst, size_t bytes) const;
  std::string getParamString(const std::string &name, const std::string &def) const;
  Object *getParamObject(const std::string &name, ANARIDataType type) const;

  // commit protocol
  virtual void commitParameters() {}
  virtual void finalize() {}
  virtual bool isValid() const { return true; }
  virtual int commitPriority() const { return 5; }
  virtual bool getProperty(const std::string &name, ANARIDataType type, void *mem, uint64_t size, uint32_t mask);
  virtual void notifyChanged(Object * /*source*/); // an observed array/object changed: re-finalise

  void addObserver(Object *o);
  void removeObserver(Object *o);
  void notifyObservers();

  void report(ANARIStatusSeverity sev, ANARIStatusCode code, const char *fmt, ...) const;

  Device *device;
  ANARIDataType type;
  std::string subtype;
  std::map<std::string, Param> params;
  uint64_t lastFinalized = 0;
  bool parametersChanged = true;

 private:
  std::atomic<int> m_public{1};
  std::atomic<int> m_internal{0};
  std::set<Object *> m_observers;
};

// intrusive INTERNAL reference
template <typename T>
struct Ref
{
  Ref() = default;
  Ref(const Ref &) = delete;
  Ref &operator=(const Ref &) = delete;
  ~Ref() { reset(); }
  void reset(T *p = nullptr)
  {
    if (p)
      p->refInc(RefType::INTERNAL);
    if (ptr)
      ptr->refDec(RefType::INTERNAL);
    ptr = p;
  }
  T *operator->() const { return ptr; }
  explicit operator bool() const { return ptr != nullptr; }
  T *ptr = nullptr;
};

// ---- arrays (devices/rtx/array/Array.cpp) ------------------------------------------------------------------
enum class Ownership
{
  SHARED,
  CAPTURED,
  MANAGED
};

struct Array : Object
{
  Array(Device *d, ANARIDataType arrayType, const void *appMemory, ANARIMemoryDeleter deleter, const void *userData,
      ANARIDataType elementType, uint64_t n1, uint64_t n2, uint64_t n3);
  ~Array() override;
  int commitPriority() const override { return 0; }
  void commitParameters() override; // Array1D "begin" / "end" region (array/Array1D.cpp:43-66)
  void *map();
  void unmap();
  void privatize(); // SHARED array lost its last public ref: copy the app memory (Array.cpp:164-182)
  const void *data() const { return m_data; }
  bool onDevice() const { return m_onDevice; }
  size_t totalSize() const { return (size_t)dims[0] * dims[1] * dims[2]; }
  size_t totalBytes() const { return totalSize() * sizeOfType(elementType); }
  Object *objectAt(size_t i) const;
  // the [begin, end) window of a 1-D array: what its consumers see (Array1D::size / begin())
  size_t regionBegin() const { return m_begin; }
  size_t regionSize() const { return m_end - m_begin; }
  const void *regionData() const { return (const uint8_t *)m_data + m_begin * sizeOfType(elementType); }
  Object *regionObjectAt(size_t i) const { return objectAt(m_begin + i); }

  ANARIDataType elementType;
  uint64_t dims[3];
  Ownership ownership;

 private:
  void *m_data = nullptr;
  bool m_onDevice = false;
  ANARIMemoryDeleter m_deleter = nullptr;
  const void *m_deleterPtr = nullptr;
  std::vector<uint8_t> m_managed;
  bool m_mapped = false;
  std::vector<Object *> m_mappedHandles; // object arrays: the elements referenced when map() was called
  size_t m_begin = 0, m_end = 0;
  ANARIDataType m_arrayType;
};

// ---- camera ---------------------------------------------------------------------------------------------------
struct Camera : Object
{
  Camera(Device *d, const std::string &subtype);
  void commitParameters() override;
  bool isValid() const override { return m_valid; }
  int commitPriority() const override { return 1; }
  DvrCamera cam{};

 private:
  bool m_valid = false;
};

// ---- spatial field ---------------------------------------------------------------------------------------------
struct SpatialField : Object
{
  SpatialField(Device *d, const std::string &subtype);
  ~SpatialField() override;
  void commitParameters() override;
  void finalize() override;
  bool isValid() const override { return m_field != nullptr || !m_parts.empty(); }
  int commitPriority() const override { return 2; }
  bool getProperty(const std::string &name, ANARIDataType type, void *mem, uint64_t size, uint32_t mask) override;
  DvrField *handle() const { return m_field; }
  // Multi-GPU: the field as GPU `rank` holds it — its z-slab in sort-last mode, a replica in sort-first mode; null when
  // the field is not distributed (single GPU, NanoVDB in sort-last mode, too few slices).
  DvrField *part(int rank) const { return (size_t)rank < m_parts.size() ? m_parts[(size_t)rank] : nullptr; }
  bool distributed() const { return !m_parts.empty(); }
  bool slabbed() const { return m_slabbed; }
  // The whole field on the display GPU, created on demand: what a frame falls back to when the scene is not one the
  // multi-GPU paths cover (several volumes in sort-last mode, dpt renderer, albedo / normal channels ...)
  DvrField *whole();
  // bumped whenever handle() becomes a NEW device field (not by an in-place refresh): an address alone can be reused
  uint64_t generation() const { return m_generation; }
  void bounds(float lo[3], float hi[3]) const;

 private:
  void cleanup();
  bool refinalizeInPlace();
  bool createDistributed(int dataType, const uint32_t dims[3], int filter);
  DvrField *createWhole(int dataType, const uint32_t dims[3], int filter);
  std::vector<DvrField *> m_parts;
  bool m_slabbed = false;
  uint64_t m_generation = 0;
  int m_fieldType = -1;
  uint32_t m_fieldDims[3] = {0, 0, 0};
  std::string m_fieldFilter;
  Ref<Array> m_data;
  float m_origin[3] = {0, 0, 0};
  float m_spacing[3] = {1, 1, 1};
  std::string m_filter = "linear";
  DvrField *m_field = nullptr;
};

// ---- volume ------------------------------------------------------------------------------------------------------
struct Volume : Object
{
  Volume(Device *d, const std::string &subtype);
  ~Volume() override;
  void commitParameters() override;
  void finalize() override;
  bool isValid() const override
  {
    return (m_volume != nullptr || !m_vparts.empty()) && m_field && m_field->isValid();
  }
  int commitPriority() const override { return 3; }
  DvrVolume *handle() const { return m_volume; }
  DvrVolume *part(int rank) const { return (size_t)rank < m_vparts.size() ? m_vparts[(size_t)rank] : nullptr; }
  bool distributed() const { return !m_vparts.empty(); }
  DvrVolume *whole(); // the volume over SpatialField::whole(), created on demand (display GPU)
  SpatialField *field() const { return m_field.ptr; }
  uint32_t id() const { return m_id; }

 private:
  void dropDeviceVolume();
  Ref<Array> m_color, m_opacity;
  Ref<SpatialField> m_field;
  float m_uniformColor[4] = {1, 1, 1, 1};
  float m_uniformOpacity = 1.f;
  float m_unitDistance = 1.f;
  float m_valueRange[2] = {0.f, 1.f};
  uint32_t m_id = ~0u;
  DvrVolume *m_volume = nullptr;
  std::vector<DvrVolume *> m_vparts;
  std::vector<float> m_tf; // the discretised table of the last finalize (for whole())
  const DvrField *m_volumeField = nullptr;
  uint64_t m_volumeFieldGeneration = 0;
  bool m_known = true;
};

// ---- surfaces and lights (what stands in front of, behind and around the volumes) ---------------------
// scene/surface/geometry/{Triangle.cpp,Sphere.cu}: the arrays stay host-side here; World::surfaceSet() hands them to
// dvr_surfaces_create, which uploads them and builds the BVH
struct Geometry : Object
{
  Geometry(Device *d, const std::string &subtype);
  void commitParameters() override;
  bool isValid() const override { return kind >= 0 && m_vertex; }
  int commitPriority() const override { return 2; }
  int kind = -1; // DvrGeometryType, -1 = a subtype this device does not render
  Array *vertex() const { return m_vertex.ptr; }
  Array *index() const { return m_index.ptr; }
  Array *normal() const { return m_normal.ptr; }
  Array *vertexRadius() const { return m_radius.ptr; }
  Array *primitiveId() const { return m_primId.ptr; }
  float radius = 0.01f;
  bool cullBackfaces = false;

 private:
  Ref<Array> m_vertex, m_index, m_normal, m_radius, m_primId;
};

// scene/surface/material/Matte.cpp:38-52 (constant colour / opacity; samplers and attribute names are not built)
struct Material : Object
{
  Material(Device *d, const std::string &subtype);
  void commitParameters() override;
  int commitPriority() const override { return 2; }
  float color[4] = {0.8f, 0.8f, 0.8f, 1.f};
  float opacity = 1.f;
  int alphaMode = DVR_ALPHA_OPAQUE;
  float alphaCutoff = 0.5f;
};

// scene/surface/Surface.cpp:42-58
struct Surface : Object
{
  Surface(Device *d);
  void commitParameters() override;
  bool isValid() const override { return m_geometry && m_geometry->isValid() && m_material; }
  int commitPriority() const override { return 3; }
  Geometry *geometry() const { return m_geometry.ptr; }
  Material *material() const { return m_material.ptr; }
  uint32_t id = ~0u;

 private:
  Ref<Geometry> m_geometry;
  Ref<Material> m_material;
};

// scene/light/{Light,Directional,Point}.cpp
struct Light : Object
{
  Light(Device *d, const std::string &subtype);
  void commitParameters() override;
  bool isValid() const override { return kind >= 0; }
  int commitPriority() const override { return 2; }
  int kind = -1; // DvrLightType
  float color[3] = {1, 1, 1};
  float vec[3] = {0, 0, -1}; // direction (normalised) or position, object space
  float strength = 1.f;
};

// ---- group / instance / world ---------------------------------------------------------------------------------------
struct Group : Object
{
  Group(Device *d);
  void commitParameters() override;
  int commitPriority() const override { return 4; }
  std::vector<Volume *> volumes() const;
  std::vector<Surface *> surfaces() const;
  std::vector<Light *> lights() const;

 private:
  Ref<Array> m_volumes, m_surfaces, m_lights;
};

struct Instance : Object
{
  Instance(Device *d, const std::string &subtype);
  void commitParameters() override;
  bool isValid() const override { return (bool)m_group; }
  int commitPriority() const override { return 5; }
  Group *group() const { return m_group.ptr; }
  float objectToWorld[12]; // column-major 4x3 (c0,c1,c2,t) as glm::mat4x3
  uint32_t id = ~0u;

 private:
  Ref<Group> m_group;
};

struct FlatInstance
{
  Volume *volume;
  float worldToObject[12]; // row-major 3x4
  uint32_t instId;
};

struct World : Object
{
  World(Device *d);
  ~World() override;
  void commitParameters() override;
  int commitPriority() const override { return 6; }
  bool getProperty(const std::string &name, ANARIDataType type, void *mem, uint64_t size, uint32_t mask) override;
  // World::rebuildWorld + the zero-instance rule (World.cpp:60-135,202-258)
  std::vector<FlatInstance> flatten(bool warn) const;
  void bounds(float lo[3], float hi[3]) const;
  // The flattened world's surfaces as a device-side set (geometry + matte material + instance transform + BVH),
  // rebuilt when anything that feeds it was finalised since the last build; null when the world has no surface
  DvrSurfaces *surfaceSet(bool warn);
  // every light instance, transformed by its instance (gpu/sampleLight.h:54-78)
  std::vector<DvrLight> flattenLights() const;

 private:
  struct FlatSurface
  {
    Surface *surface;
    float objectToWorld[12]; // row-major 3x4
    uint32_t instId;
  };
  std::vector<FlatSurface> flattenSurfaces(bool warn) const;
  void dropSurfaceSet();
  Ref<Array> m_zeroVolumes, m_instances, m_zeroSurfaces, m_zeroLights;
  DvrSurfaces *m_surfaceSet = nullptr;
  uint64_t m_surfaceStamp = 0; // fingerprint of what m_surfaceSet was built from
  int m_surfaceSetGpu = -1;
};

// ---- renderer ----------------------------------------------------------------------------------------------------
struct Renderer : Object
{
  Renderer(Device *d, const std::string &subtype);
  ~Renderer() override;
  void commitParameters() override;
  void finalize() override; // background image -> RGBA8 texture (Renderer.cpp:172-179)
  bool isValid() const override { return m_known; }
  int commitPriority() const override { return 1; }
  float background[4] = {0, 0, 0, 1};
  const DvrImage *backgroundImage() const { return m_bgImage; }
  int spp = 1;
  int sampleLimit = 128;
  bool checkerboard = false;
  float volumeSamplingRate = 0.125f;
  int integrator = DVR_INTEGRATOR_DEFAULT;
  int maxDepth = 5;                // dpt only
  bool dptReferenceGrid = false;   // dpt only (extension)
  float ambientRadiance = 0.f;     // dpt: light of the walk; default / directLight: ambient term of surface shading
  float occlusionDistance = 1e20f; // dpt: scatter ray length; default / directLight: ambient-occlusion ray length
  float ambientColor[3] = {1, 1, 1}; // surface shading only (Renderer.cpp:158)
  int ambientSamples = 1;            // "ambientSamples" (DirectLight.cpp:52)
  bool cullTriangleBackfaces = false;
  int macrocellSkipping = DVR_SKIP_AUTO;
  // sort-first extension: this device renders only tile rows (row % tileRanks == tileRank)
  uint32_t tileRank = 0, tileRanks = 1;

 private:
  void dropBackgroundImage();
  bool m_known = true;
  Ref<Array> m_bgArray; // "background" given as an Array2D
  DvrImage *m_bgImage = nullptr;
};

// ---- frame ---------------------------------------------------------------------------------------------------------
struct Frame : Object
{
  Frame(Device *d);
  ~Frame() override;
  void commitParameters() override;
  void finalize() override;
  bool isValid() const override;
  int commitPriority() const override { return 7; }
  bool getProperty(const std::string &name, ANARIDataType type, void *mem, uint64_t size, uint32_t mask) override;

  void renderFrame();
  const void *map(const std::string &channel, uint32_t *w, uint32_t *h, ANARIDataType *type);
  int ready(ANARIWaitMask m);
  void wait() const;

 private:
  void checkAccumulationReset();
  void freeBuffers();
  // multi-GPU frames (device.cpp): per-GPU accumulation / partial images / flag tables, fused slab frames
  struct PerGpu
  {
    void *accum = nullptr, *depth = nullptr;        // this GPU's share of the accumulation state
    void *partial[2] = {nullptr, nullptr};          // float4[W*H] + float[W*H], alternating between frames
    unsigned int *flags = nullptr;                  // region table + resolved table + error word
    unsigned int *regionDone = nullptr;
    void *done = nullptr;                           // sort-first: "this GPU's tile rows are rendered" event
  };
  std::vector<PerGpu> m_perGpu;
  uint32_t m_seq = 0;
  bool ensureMultiGpuBuffers();
  void freeMultiGpuBuffers();
  void syncAllGpus() const;
  bool renderSortLast(const DvrFrameParams &p, const DvrFrameBuffers &display, Volume *v, const FlatInstance &fi);
  bool renderSortFirst(DvrFrameParams p, const DvrFrameBuffers &display, const std::vector<FlatInstance> &flat);
  void *download(void *dev, size_t bytes, std::vector<uint8_t> &host);

  Ref<Renderer> m_renderer;
  Ref<Camera> m_camera;
  Ref<World> m_world;
  ANARIFrameCompletionCallback m_callback = nullptr;
  const void *m_callbackUserPtr = nullptr;
  ANARIDataType m_colorType = ANARI_UFIXED8_RGBA_SRGB, m_depthType = ANARI_UNKNOWN, m_primIdType = ANARI_UNKNOWN,
                m_objIdType = ANARI_UNKNOWN, m_instIdType = ANARI_UNKNOWN, m_albedoType = ANARI_UNKNOWN,
                m_normalType = ANARI_UNKNOWN;
  uint32_t m_size[2] = {10, 10};
  int m_format = DVR_FORMAT_UFIXED8_RGBA_SRGB;
  // device buffers
  void *m_accum = nullptr, *m_color = nullptr, *m_depth = nullptr, *m_primId = nullptr, *m_objId = nullptr,
       *m_instId = nullptr, *m_albedoAccum = nullptr, *m_normalAccum = nullptr, *m_albedo = nullptr,
       *m_normal = nullptr;
  std::vector<uint8_t> m_hColor, m_hDepth, m_hPrim, m_hObj, m_hInst, m_hAlbedo, m_hNormal;
  void *m_pinned = nullptr; // pinned host copy of the colour channel (device-accessible under UVA)
  size_t m_pinnedBytes = 0;
  // Once the application has mapped channel.color to the host, later launches also stream the encoded colour
  // straight into m_pinned (DvrFrameBuffers::outColorMirror), so the next map needs no copy after the march.
  bool m_streamColorToHost = false;
  bool m_pinnedHoldsFrame = false;
  bool ensurePinned(size_t bytes);
  void *m_eventStart = nullptr, *m_eventEnd = nullptr;
  int m_frameID = 0, m_checkerboardID = -1;
  float m_invFrameID = 1.f;
  bool m_nextFrameReset = true;
  bool m_valid = false;
  bool m_everRendered = false;
  uint64_t m_lastCommitSeen = 0;
  float m_duration = 0.f;
};

// ---- device -----------------------------------------------------------------------------------------------------------
struct Device : Object
{
  Device(ANARIStatusCallback cb, const void *userPtr);
  ~Device() override;
  void commitParameters() override; // statusCallback, cudaDevice, forceInit
  bool getProperty(const std::string &name, ANARIDataType type, void *mem, uint64_t size, uint32_t mask) override;

  bool initDevice(); // lazy CUDA init, sticky failure (VisRTXDevice.cpp:435-458)
  void enqueueCommit(Object *o);
  void flushCommits();
  void removeFromQueue(Object *o);
  uint64_t newTimeStamp() { return ++m_clock; }
  uint64_t lastFinalization() const { return m_lastFinalization; }
  void *stream() const { return m_stream; }
  int cudaDevice() const { return m_gpuID; }
  // Multi-GPU (one process, peer access; the reference is single-GPU, VisRTXDevice.cpp:464): the device parameter
  // "cudaDevices" lists the GPUs, the first one being the display GPU that owns the frame's output channels;
  // "multiGpuMode" = "sortLast" (z-slabs of every structuredRegular field, fused march + exchange per GPU, default)
  // or "sortFirst" (fields replicated, interleaved tile rows).
  int gpuCount() const { return (int)m_gpus.size(); }
  int gpu(int rank) const { return m_gpus[(size_t)rank]; }
  void *stream(int rank) const { return m_streams[(size_t)rank]; }
  bool sortLast() const { return gpuCount() > 1 && m_sortLast; }
  bool sortFirst() const { return gpuCount() > 1 && !m_sortLast; }

  void message(const Object *src, ANARIStatusSeverity sev, ANARIStatusCode code, const char *msg) const;

  std::recursive_mutex mutex;

 private:
  ANARIStatusCallback m_cb = nullptr;
  const void *m_cbUserPtr = nullptr;
  ANARIStatusCallback m_defaultCb = nullptr;
  const void *m_defaultCbUserPtr = nullptr;
  std::vector<Object *> m_commitQueue;
  std::atomic<uint64_t> m_clock{0};
  uint64_t m_lastFinalization = 0;
  int m_initStatus = 0; // 0 uninitialised, 1 ok, -1 failed
  int m_desiredGpuID = 0, m_gpuID = -1;
  void *m_stream = nullptr;
  std::vector<int> m_desiredGpus; // "cudaDevices"; empty = {cudaDevice}
  std::vector<int> m_gpus;        // after initDevice: [0] == m_gpuID
  std::vector<void *> m_streams;  // one private stream per GPU, [0] == m_stream
  bool m_sortLast = true;
};

// RAII: cudaSetDevice(gpu) for the scope
struct GpuScope
{
  explicit GpuScope(int gpu);
  ~GpuScope();
  int prev = -1;
};

// RAII: every API entry saves / restores the caller's current CUDA device (VisRTXDevice.cpp:790-814)
struct CudaDeviceScope
{
  explicit CudaDeviceScope(Device *d);
  ~CudaDeviceScope();
  int prev = -1;
  bool active = false;
};

// ---- introspection (queries.cpp) -----------------------------------------------------------------------------------
const char **querySubtypes(ANARIDataType type);
const void *queryObjectInfo(ANARIDataType type, const char *subtype, const char *infoName, ANARIDataType infoType,
    const char **extensions);
const void *queryParameterInfo(ANARIDataType type, const char *subtype, const char *paramName, ANARIDataType paramType,
    const char *infoName, ANARIDataType infoType);

} // namespace b200
