// objects.cpp — object model implementation (see objects.h).  No arithmetic of the hot path lives
// here: fields/volumes/cameras are translated to the C-ABI of include/dvr_b200.h.
#include "objects.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <limits>

namespace b200 {

// ---------------------------------------------------------------------------------------------------------
size_t sizeOfType(ANARIDataType t)
{
  if (t >= 1000 && t < 1076) { // scalar/vector families of four
    static const size_t base[] = {1, 1, 2, 2, 4, 4, 8, 8, 1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8};
    const int fam = (t - 1000) / 4, n = (t - 1000) % 4 + 1;
    return base[fam] * n;
  }
  switch (t) {
  case ANARI_BOOL: return 4; // ANARI bools are int32 on the API
  case ANARI_DATA_TYPE: return sizeof(ANARIDataType);
  case ANARI_STRING: return sizeof(char *);
  case ANARI_VOID_POINTER:
  case ANARI_FUNCTION_POINTER:
  case ANARI_MEMORY_DELETER:
  case ANARI_STATUS_CALLBACK:
  case ANARI_FRAME_COMPLETION_CALLBACK: return sizeof(void *);
  case ANARI_UFIXED8_R_SRGB: return 1;
  case ANARI_UFIXED8_RA_SRGB: return 2;
  case ANARI_UFIXED8_RGB_SRGB: return 3;
  case ANARI_UFIXED8_RGBA_SRGB: return 4;
  case ANARI_INT32_BOX1: return 8;
  case ANARI_FLOAT32_BOX1: return 8;
  case ANARI_FLOAT32_BOX2: return 16;
  case ANARI_FLOAT32_BOX3: return 24;
  case ANARI_FLOAT32_BOX4: return 32;
  case ANARI_FLOAT32_MAT2: return 16;
  case ANARI_FLOAT32_MAT3: return 36;
  case ANARI_FLOAT32_MAT4: return 64;
  case ANARI_FLOAT32_MAT2x3: return 24;
  case ANARI_FLOAT32_MAT3x4: return 48;
  case ANARI_FLOAT32_QUAT_IJKW: return 16;
  case ANARI_UINT64_REGION1: return 16;
  case ANARI_FLOAT64_BOX1: return 16;
  default: break;
  }
  if (isObjectType(t))
    return sizeof(void *);
  return 0;
}

bool isObjectType(ANARIDataType t) { return t >= ANARI_LIBRARY && t <= ANARI_WORLD; }

const char *typeName(ANARIDataType t)
{
  switch (t) {
  case ANARI_FLOAT32: return "ANARI_FLOAT32";
  case ANARI_FLOAT64: return "ANARI_FLOAT64";
  case ANARI_FIXED8: return "ANARI_FIXED8";
  case ANARI_UFIXED8: return "ANARI_UFIXED8";
  case ANARI_FIXED16: return "ANARI_FIXED16";
  case ANARI_UFIXED16: return "ANARI_UFIXED16";
  case ANARI_FLOAT16: return "ANARI_FLOAT16";
  case ANARI_UINT8: return "ANARI_UINT8";
  case ANARI_UINT32: return "ANARI_UINT32";
  case ANARI_FLOAT32_VEC3: return "ANARI_FLOAT32_VEC3";
  case ANARI_FLOAT32_VEC4: return "ANARI_FLOAT32_VEC4";
  case ANARI_VOLUME: return "ANARI_VOLUME";
  case ANARI_INSTANCE: return "ANARI_INSTANCE";
  default: return "ANARI_<type>";
  }
}

// ---------------------------------------------------------------------------------------------------------
// Object
// ---------------------------------------------------------------------------------------------------------
Object::Object(Device *d, ANARIDataType t, std::string st) : device(d), type(t), subtype(std::move(st)) {}

Object::~Object()
{
  for (auto &kv : params)
    if (kv.second.object)
      kv.second.object->refDec(RefType::INTERNAL);
  if (device && device != this)
    device->removeFromQueue(this);
}

void Object::refInc(RefType t) { (t == RefType::PUBLIC ? m_public : m_internal)++; }

void Object::refDec(RefType t)
{
  auto &c = (t == RefType::PUBLIC ? m_public : m_internal);
  if (c.load() > 0)
    c--;
  if (t == RefType::PUBLIC && m_public.load() == 0 && type >= ANARI_ARRAY && type <= ANARI_ARRAY3D)
    static_cast<Array *>(this)->privatize();
  if (m_public.load() == 0 && m_internal.load() == 0)
    delete this;
}

void Object::setParam(const char *name, ANARIDataType t, const void *mem)
{
  if (!name || !mem)
    return;
  Param p;
  p.type = t;
  if (t == ANARI_STRING) {
    const char *s = (const char *)mem;
    p.bytes.assign(s, s + std::strlen(s) + 1);
  } else if (isObjectType(t)) {
    Object *o = *(Object *const *)mem;
    p.object = o;
    if (o)
      o->refInc(RefType::INTERNAL);
  } else {
    const size_t n = sizeOfType(t);
    if (n == 0) {
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "parameter '%s' has an unknown data type (%d)",
          name, t);
      return;
    }
    p.bytes.assign((const uint8_t *)mem, (const uint8_t *)mem + n);
  }
  auto it = params.find(name);
  if (it != params.end() && it->second.object)
    it->second.object->refDec(RefType::INTERNAL);
  params[name] = std::move(p);
  parametersChanged = true;
}

void Object::unsetParam(const char *name)
{
  auto it = params.find(name);
  if (it == params.end())
    return;
  if (it->second.object)
    it->second.object->refDec(RefType::INTERNAL);
  params.erase(it);
  parametersChanged = true;
}

void Object::unsetAllParams()
{
  for (auto &kv : params)
    if (kv.second.object)
      kv.second.object->refDec(RefType::INTERNAL);
  params.clear();
  parametersChanged = true;
}

const Param *Object::findParam(const std::string &name) const
{
  auto it = params.find(name);
  return it == params.end() ? nullptr : &it->second;
}

bool Object::getParamRaw(const std::string &name, ANARIDataType t, void *dst, size_t bytes) const
{
  const Param *p = findParam(name);
  if (!p || p->type != t || p->bytes.size() < bytes)
    return false;
  std::memcpy(dst, p->bytes.data(), bytes);
  return true;
}

std::string Object::getParamString(const std::string &name, const std::string &def) const
{
  const Param *p = findParam(name);
  if (!p || p->type != ANARI_STRING || p->bytes.empty())
    return def;
  return std::string((const char *)p->bytes.data());
}

Object *Object::getParamObject(const std::string &name, ANARIDataType t) const
{
  const Param *p = findParam(name);
  if (!p || !p->object)
    return nullptr;
  // ANARI_ARRAY matches any array rank; exact match otherwise
  if (p->object->type == t)
    return p->object;
  if (t == ANARI_ARRAY && p->object->type >= ANARI_ARRAY1D && p->object->type <= ANARI_ARRAY3D)
    return p->object;
  return nullptr;
}

bool Object::getProperty(const std::string &, ANARIDataType, void *, uint64_t, uint32_t) { return false; }

void Object::notifyChanged(Object *)
{
  if (device)
    device->enqueueCommit(this);
}

void Object::addObserver(Object *o) { m_observers.insert(o); }
void Object::removeObserver(Object *o) { m_observers.erase(o); }
void Object::notifyObservers()
{
  const auto obs = m_observers; // observers may re-register while being notified
  for (Object *o : obs)
    o->notifyChanged(this);
}

void Object::report(ANARIStatusSeverity sev, ANARIStatusCode code, const char *fmt, ...) const
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  Device *d = device ? device : (Device *)this;
  d->message(this, sev, code, buf);
}

// ---------------------------------------------------------------------------------------------------------
// Array
// ---------------------------------------------------------------------------------------------------------
static bool pointerIsDevice(const void *p)
{ // array/Array.cpp:38-66
  if (!p)
    return false;
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

Array::Array(Device *d, ANARIDataType arrayType, const void *appMemory, ANARIMemoryDeleter deleter,
    const void *userData, ANARIDataType et, uint64_t n1, uint64_t n2, uint64_t n3)
    : Object(d, arrayType), elementType(et), m_arrayType(arrayType)
{
  dims[0] = n1;
  dims[1] = n2;
  dims[2] = n3;
  m_begin = 0;
  m_end = (size_t)(n1 * n2 * n3);
  if (appMemory) {
    ownership = deleter ? Ownership::CAPTURED : Ownership::SHARED;
    m_data = const_cast<void *>(appMemory);
    m_onDevice = pointerIsDevice(appMemory);
    m_deleter = deleter;
    m_deleterPtr = userData;
    if (isObjectType(et) && m_onDevice) {
      report(ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_OPERATION,
          "illegal operation: cannot create object arrays from GPU memory");
      m_data = nullptr;
    }
  } else {
    ownership = Ownership::MANAGED;
    m_managed.assign(totalBytes(), 0); // zero-initialised (Array.cpp:113-116)
    m_data = m_managed.data();
  }
  if (isObjectType(et) && m_data && !m_onDevice) {
    // object arrays hold INTERNAL refs on their elements for as long as they reference them
    for (size_t i = 0; i < totalSize(); ++i)
      if (Object *o = ((Object **)m_data)[i])
        o->refInc(RefType::INTERNAL);
  }
}

Array::~Array()
{
  if (isObjectType(elementType) && m_data && !m_onDevice)
    for (size_t i = 0; i < totalSize(); ++i)
      if (Object *o = ((Object **)m_data)[i])
        o->refDec(RefType::INTERNAL);
  if (ownership == Ownership::CAPTURED && m_deleter)
    m_deleter(m_deleterPtr, m_data);
}

Object *Array::objectAt(size_t i) const
{
  if (!isObjectType(elementType) || !m_data || i >= totalSize())
    return nullptr;
  return ((Object **)m_data)[i];
}

void *Array::map()
{
  if (m_mapped)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_OPERATION, "array mapped again without unmapping");
  m_mapped = true;
  // Elements may be replaced while mapped.  The references held on the current elements are KEPT for the duration of
  // the map (an element whose only owner is this array must stay alive: World / Group still read it); the handle list
  // is remembered and reconciled with the new contents at unmap, as helium's ObjectArray does.
  m_mappedHandles.clear();
  if (isObjectType(elementType) && m_data && !m_onDevice)
    m_mappedHandles.assign((Object **)m_data, (Object **)m_data + totalSize());
  return m_data;
}

void Array::commitParameters()
{ // Array1D::commitParameters / finalize, array/Array1D.cpp:43-66
  if (m_arrayType != ANARI_ARRAY1D)
    return;
  const size_t capacity = totalSize();
  if (capacity == 0)
    return;
  size_t b = (size_t)getParam<uint64_t>("begin", ANARI_UINT64, 0);
  size_t e = (size_t)getParam<uint64_t>("end", ANARI_UINT64, (uint64_t)capacity);
  b = std::min(std::max(b, (size_t)0), capacity - 1);
  e = std::min(std::max(e, (size_t)1), capacity);
  if (b > e) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "array 'begin' is not less than 'end', swapping values");
    std::swap(b, e);
  }
  if (b != m_begin || e != m_end) {
    m_begin = b;
    m_end = e;
    notifyObservers(); // markDataModified + notifyChangeObservers
  }
}

void Array::unmap()
{
  if (!m_mapped) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_OPERATION, "array unmapped again without mapping");
    return;
  }
  m_mapped = false;
  if (isObjectType(elementType) && m_data && !m_onDevice) {
    for (size_t i = 0; i < totalSize(); ++i) // new contents first, so an element present before and after never hits 0
      if (Object *o = ((Object **)m_data)[i])
        o->refInc(RefType::INTERNAL);
    for (Object *o : m_mappedHandles)
      if (o)
        o->refDec(RefType::INTERNAL);
    m_mappedHandles.clear();
  }
  notifyObservers(); // Array.cpp:152-162: data modified => observing field/volume re-finalises
}

void Array::privatize()
{
  if (ownership != Ownership::SHARED || !m_data || useCount(RefType::INTERNAL) == 0)
    return;
  const size_t bytes = totalBytes();
  if (m_onDevice) {
    void *copy = nullptr;
    if (cudaMalloc(&copy, bytes) == cudaSuccess) {
      cudaMemcpy(copy, m_data, bytes, cudaMemcpyDeviceToDevice);
      m_data = copy;
      ownership = Ownership::CAPTURED;
      m_deleter = [](const void *, const void *mem) { cudaFree(const_cast<void *>(mem)); };
      m_deleterPtr = nullptr;
    }
  } else {
    m_managed.assign((const uint8_t *)m_data, (const uint8_t *)m_data + bytes);
    m_data = m_managed.data();
    ownership = Ownership::MANAGED;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Camera (camera/Camera.cpp:68-76, Perspective.cpp:42-72, Orthographic.cpp:38-52)
// ---------------------------------------------------------------------------------------------------------
Camera::Camera(Device *d, const std::string &st) : Object(d, ANARI_CAMERA, st)
{
  d->enqueueCommit(this); // cameras commit their defaults even without an explicit commit (Camera.cpp:42-46)
}

void Camera::commitParameters()
{
  float region[4] = {0.f, 0.f, 1.f, 1.f};
  getParamRaw("imageRegion", ANARI_FLOAT32_BOX2, region, sizeof(region));
  float pos[3] = {0, 0, 0}, dir[3] = {0, 0, 1}, up[3] = {0, 1, 0};
  getParamRaw("position", ANARI_FLOAT32_VEC3, pos, sizeof(pos));
  getParamRaw("direction", ANARI_FLOAT32_VEC3, dir, sizeof(dir));
  getParamRaw("up", ANARI_FLOAT32_VEC3, up, sizeof(up));
  const float aspect = getParam<float>("aspect", ANARI_FLOAT32, 1.f);
  m_valid = true;
  if (subtype == "perspective") {
    const float fovy = getParam<float>("fovy", ANARI_FLOAT32, 60.f * 3.14159265358979323846f / 180.f);
    const float focus = getParam<float>("focusDistance", ANARI_FLOAT32, 1.f);
    const float aperture = getParam<float>("apertureRadius", ANARI_FLOAT32, 0.f);
    dvr_camera_perspective(pos, dir, up, fovy, aspect, focus, aperture, region, &cam);
  } else if (subtype == "orthographic") {
    const float height = getParam<float>("height", ANARI_FLOAT32, 1.f);
    dvr_camera_orthographic(pos, dir, up, height, aspect, region, &cam);
  } else {
    m_valid = false;
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "unknown camera subtype '%s'", subtype.c_str());
  }
}

// ---------------------------------------------------------------------------------------------------------
// SpatialField: structuredRegular (spatial_field/StructuredRegularField.cpp:90-194)
// ---------------------------------------------------------------------------------------------------------
SpatialField::SpatialField(Device *d, const std::string &st) : Object(d, ANARI_SPATIAL_FIELD, st) {}
SpatialField::~SpatialField()
{
  if (m_data)
    m_data->removeObserver(this);
  cleanup();
}

// Generations are drawn from ONE counter for all fields: a new DvrField may reuse the heap address of a destroyed
// one, so (address, per-object counter) pairs can collide across SpatialField objects — a global stamp cannot.
static std::atomic<uint64_t> g_fieldGeneration{0};

void SpatialField::cleanup()
{
  for (size_t r = 0; r < m_parts.size(); ++r) {
    if (!m_parts[r] || m_parts[r] == m_field)
      continue;
    GpuScope scope(device->gpu((int)r));
    cudaStreamSynchronize((cudaStream_t)device->stream((int)r));
    dvr_field_destroy(m_parts[r]);
  }
  m_parts.clear();
  m_slabbed = false;
  if (m_field) {
    CudaDeviceScope scope(device);
    cudaStreamSynchronize((cudaStream_t)device->stream());
    dvr_field_destroy(m_field);
    m_field = nullptr;
    m_fieldType = -1;
  }
  m_generation = ++g_fieldGeneration;
}

// Balanced, contiguous z-slice ownership [z0, z1) of `rank` among `world` GPUs (every slice owned exactly once)
static void slabRange(uint32_t nz, int world, int rank, uint32_t &z0, uint32_t &z1)
{
  const uint32_t base = nz / (uint32_t)world, rem = nz % (uint32_t)world;
  z0 = (uint32_t)rank * base + std::min((uint32_t)rank, rem);
  z1 = z0 + base + ((uint32_t)rank < rem ? 1u : 0u);
}

// Multi-GPU: the field on every GPU of the device.  Sort-last: z-slabs (own slices + one ghost slice each side,
// dvr_field_create_structured_slab) — the volume may be larger than one GPU's memory; sort-first: a replica per GPU.
bool SpatialField::createDistributed(int dataType, const uint32_t dims[3], int filter)
{
  const int world = device->gpuCount();
  const size_t esz = sizeOfType(m_data->elementType);
  const bool slabs = device->sortLast() && dims[2] >= (uint32_t)(2 * world);
  if (device->sortLast() && !slabs)
    return false; // too thin to slice: the caller keeps it whole on the display GPU
  m_parts.assign((size_t)world, nullptr);
  for (int r = 0; r < world; ++r) {
    GpuScope scope(device->gpu(r));
    int rc;
    if (slabs) {
      uint32_t z0, z1;
      slabRange(dims[2], world, r, z0, z1);
      const uint32_t first = z0 > 0 ? z0 - 1 : 0; // data points at the first RESIDENT slice (ghost included)
      const uint8_t *src = (const uint8_t *)m_data->data() + (size_t)first * dims[0] * dims[1] * esz;
      rc = dvr_field_create_structured_slab(src, m_data->onDevice() ? 1 : 0, dataType, dims, z0, z1, m_origin, m_spacing,
          filter, device->stream(r), &m_parts[(size_t)r]);
    } else
      rc = dvr_field_create_structured(m_data->data(), m_data->onDevice() ? 1 : 0, dataType, dims, m_origin, m_spacing,
          filter, device->stream(r), &m_parts[(size_t)r]);
    if (rc != DVR_OK) {
      m_parts[(size_t)r] = nullptr;
      report(ANARI_SEVERITY_ERROR, rc == DVR_ERR_OUT_OF_MEMORY ? ANARI_STATUS_OUT_OF_MEMORY : ANARI_STATUS_UNKNOWN_ERROR,
          "structuredRegular field upload to GPU %d failed: %s", device->gpu(r), dvr_last_error());
      cleanup();
      return false;
    }
  }
  m_slabbed = slabs;
  if (!slabs)
    m_field = m_parts[0]; // the display GPU's replica doubles as the whole field
  return true;
}

DvrField *SpatialField::createWhole(int dataType, const uint32_t dims[3], int filter)
{
  CudaDeviceScope scope(device);
  DvrField *f = nullptr;
  const int rc = dvr_field_create_structured(m_data->data(), m_data->onDevice() ? 1 : 0, dataType, dims, m_origin,
      m_spacing, filter, device->stream(), &f);
  if (rc != DVR_OK) {
    report(ANARI_SEVERITY_ERROR, rc == DVR_ERR_OUT_OF_MEMORY ? ANARI_STATUS_OUT_OF_MEMORY : ANARI_STATUS_UNKNOWN_ERROR,
        "structuredRegular field upload failed: %s", dvr_last_error());
    return nullptr;
  }
  return f;
}

DvrField *SpatialField::whole()
{
  if (m_field || !m_slabbed || !m_data)
    return m_field;
  report(ANARI_SEVERITY_PERFORMANCE_WARNING, ANARI_STATUS_NO_ERROR,
      "this frame is outside what the sort-last path covers: uploading the whole field to the display GPU as well");
  m_field = createWhole(m_fieldType, m_fieldDims, m_fieldFilter == "nearest" ? DVR_FILTER_NEAREST : DVR_FILTER_LINEAR);
  return m_field;
}

void SpatialField::commitParameters()
{
  m_origin[0] = m_origin[1] = m_origin[2] = 0.f;
  m_spacing[0] = m_spacing[1] = m_spacing[2] = 1.f;
  getParamRaw("origin", ANARI_FLOAT32_VEC3, m_origin, sizeof(m_origin));
  getParamRaw("spacing", ANARI_FLOAT32_VEC3, m_spacing, sizeof(m_spacing));
  m_filter = getParamString("filter", "linear");
  Array *a = static_cast<Array *>(getParamObject("data", subtype == "nanovdb" ? ANARI_ARRAY1D : ANARI_ARRAY3D));
  if (m_data.ptr != a) {
    if (m_data)
      m_data->removeObserver(this);
    m_data.reset(a);
    if (a)
      a->addObserver(this);
  }
}

static int dvrTypeOf(ANARIDataType t)
{
  switch (t) {
  case ANARI_FLOAT32: return DVR_FLOAT32;
  case ANARI_UFIXED8: return DVR_UFIXED8;
  case ANARI_FIXED8: return DVR_FIXED8;
  case ANARI_UFIXED16: return DVR_UFIXED16;
  case ANARI_FIXED16: return DVR_FIXED16;
  case ANARI_FLOAT64: return DVR_FLOAT64;
  case ANARI_FLOAT16: return DVR_FLOAT16; // extension over the reference (BASELINE config 3)
  default: return -1;
  }
}

void SpatialField::finalize()
{
  if (subtype == "structuredRegular" && m_field && m_data && refinalizeInPlace())
    return;
  cleanup();
  if (subtype == "nanovdb") { // spatial_field/NvdbRegularField.cpp:64-113
    if (!m_data) {
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
          "missing required parameter 'data' on NanoVDB regular spatial field");
      return;
    }
    if (m_data->elementType != ANARI_UINT8) {
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
          "invalid data array type encountered in NanoVDB spatial field(%s)", typeName(m_data->elementType));
      return;
    }
    if (!device->initDevice())
      return;
    CudaDeviceScope scope(device);
    const int rc = dvr_field_create_nanovdb(m_data->data(), m_data->totalBytes(), m_data->onDevice() ? 1 : 0,
        device->stream(), &m_field);
    if (rc != DVR_OK) {
      m_field = nullptr;
      report(rc == DVR_ERR_UNSUPPORTED ? ANARI_SEVERITY_WARNING : ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_ARGUMENT,
          "NanoVDB field rejected: %s", dvr_last_error());
    }
    return;
  }
  if (subtype != "structuredRegular") {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "unknown spatial field subtype '%s'",
        subtype.c_str());
    return;
  }
  if (!m_data) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "missing required parameter 'data' on structuredRegular spatial field");
    return;
  }
  const int dt = dvrTypeOf(m_data->elementType);
  if (dt < 0) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "invalid data array type encountered in structuredRegular spatial field(%s)", typeName(m_data->elementType));
    return;
  }
  if (!device->initDevice())
    return;
  CudaDeviceScope scope(device);
  const uint32_t dims[3] = {(uint32_t)m_data->dims[0], (uint32_t)m_data->dims[1], (uint32_t)m_data->dims[2]};
  const int filt = m_filter == "nearest" ? DVR_FILTER_NEAREST : DVR_FILTER_LINEAR;
  if (!(device->gpuCount() > 1 && dt != DVR_FLOAT64 && createDistributed(dt, dims, filt))) {
    m_field = createWhole(dt, dims, filt);
    if (!m_field)
      return;
  }
  m_fieldType = dt;
  m_fieldFilter = m_filter;
  for (int i = 0; i < 3; ++i)
    m_fieldDims[i] = dims[i];
}

// Time-varying fields (the `data` array rewritten and the field re-committed every step): when shape, element type
// and filter are unchanged the device array, textures and macrocell storage are reused instead of rebuilt.
bool SpatialField::refinalizeInPlace()
{
  if (distributed())
    return false; // slabs / replicas are rebuilt
  const int dt = dvrTypeOf(m_data->elementType);
  if (dt < 0 || dt != m_fieldType || m_filter != m_fieldFilter)
    return false;
  for (int i = 0; i < 3; ++i)
    if ((uint32_t)m_data->dims[i] != m_fieldDims[i])
      return false;
  CudaDeviceScope scope(device);
  return dvr_field_update_structured(m_field, m_data->data(), m_data->onDevice() ? 1 : 0, dt, m_origin, m_spacing,
             device->stream())
      == DVR_OK;
}

void SpatialField::bounds(float lo[3], float hi[3]) const
{
  if (m_field)
    dvr_field_bounds(m_field, lo, hi);
  else if (!m_parts.empty() && m_parts[0])
    dvr_field_bounds(m_parts[0], lo, hi); // a slab reports the bounds of the whole field
  else {
    lo[0] = lo[1] = lo[2] = 0.f;
    hi[0] = hi[1] = hi[2] = 1.f;
  }
}

bool SpatialField::getProperty(const std::string &name, ANARIDataType t, void *mem, uint64_t size, uint32_t mask)
{
  if (name == "bounds" && t == ANARI_FLOAT32_BOX3 && size >= 24) {
    if (mask & ANARI_WAIT)
      device->flushCommits();
    float b[6];
    bounds(b, b + 3);
    std::memcpy(mem, b, 24);
    return true;
  }
  if (name == "valueRange" && t == ANARI_FLOAT32_BOX1 && size >= 8) { // tsd computeScalarRange equivalent
    if (mask & ANARI_WAIT)
      device->flushCommits();
    if (!isValid())
      return false;
    float r[2] = {std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
    if (m_field) {
      CudaDeviceScope scope(device);
      if (dvr_field_value_range(m_field, device->stream(), r) != DVR_OK)
        return false;
    } else
      for (size_t k = 0; k < m_parts.size(); ++k) { // slabs: the extrema of all of them
        GpuScope scope(device->gpu((int)k));
        float q[2];
        if (dvr_field_value_range(m_parts[k], device->stream((int)k), q) != DVR_OK)
          return false;
        r[0] = std::min(r[0], q[0]);
        r[1] = std::max(r[1], q[1]);
      }
    std::memcpy(mem, r, 8);
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------
// Volume: transferFunction1D | scivis (scene/volume/Volume.cpp:40-90, TransferFunction1D.cpp:46-150)
// ---------------------------------------------------------------------------------------------------------
Volume::Volume(Device *d, const std::string &st) : Object(d, ANARI_VOLUME, st)
{
  m_known = (st == "transferFunction1D" || st == "scivis");
}

Volume::~Volume()
{
  if (m_color) m_color->removeObserver(this);
  if (m_opacity) m_opacity->removeObserver(this);
  if (m_field) m_field->removeObserver(this);
  dropDeviceVolume();
}

void Volume::commitParameters()
{
  m_id = getParam<uint32_t>("id", ANARI_UINT32, ~0u);
  auto observe = [this](Ref<Array> &slot, Array *a) {
    if (slot.ptr == a)
      return;
    if (slot)
      slot->removeObserver(this);
    slot.reset(a);
    if (a)
      a->addObserver(this);
  };
  observe(m_color, static_cast<Array *>(getParamObject("color", ANARI_ARRAY1D)));
  observe(m_opacity, static_cast<Array *>(getParamObject("opacity", ANARI_ARRAY1D)));
  m_uniformColor[0] = m_uniformColor[1] = m_uniformColor[2] = m_uniformColor[3] = 1.f;
  getParamRaw("color", ANARI_FLOAT32_VEC3, m_uniformColor, 12);
  getParamRaw("color", ANARI_FLOAT32_VEC4, m_uniformColor, 16);
  m_uniformOpacity = getParam<float>("opacity", ANARI_FLOAT32, 1.f) * m_uniformColor[3];
  m_unitDistance = getParam<float>("unitDistance", ANARI_FLOAT32, 1.f);
  SpatialField *f = static_cast<SpatialField *>(getParamObject("value", ANARI_SPATIAL_FIELD));
  if (m_field.ptr != f) {
    if (m_field)
      m_field->removeObserver(this);
    m_field.reset(f);
    if (f)
      f->addObserver(this);
  }
  m_valueRange[0] = 0.f;
  m_valueRange[1] = 1.f;
  getParamRaw("valueRange", ANARI_FLOAT32_VEC2, m_valueRange, 8);
  getParamRaw("valueRange", ANARI_FLOAT32_BOX1, m_valueRange, 8);
  double vr[2];
  if (getParamRaw("valueRange", ANARI_FLOAT64_BOX1, vr, 16)) {
    m_valueRange[0] = float(vr[0]);
    m_valueRange[1] = float(vr[1]);
  }
}

void Volume::dropDeviceVolume()
{
  for (size_t r = 0; r < m_vparts.size(); ++r) {
    if (!m_vparts[r] || m_vparts[r] == m_volume)
      continue;
    GpuScope scope(device->gpu((int)r));
    cudaStreamSynchronize((cudaStream_t)device->stream((int)r));
    dvr_volume_destroy(m_vparts[r]);
  }
  m_vparts.clear();
  if (!m_volume)
    return;
  CudaDeviceScope scope(device);
  cudaStreamSynchronize((cudaStream_t)device->stream());
  dvr_volume_destroy(m_volume);
  m_volume = nullptr;
  m_volumeField = nullptr;
}

void Volume::finalize()
{
  // The device volume references its DvrField by pointer: once that field was destroyed or replaced the volume must
  // not survive an early return below (isValid() would otherwise vouch for a dangling handle).
  if ((m_volume || !m_vparts.empty())
      && (!m_field || !m_field->isValid() || m_field->distributed() || m_volumeField != m_field->handle()
          || m_volumeFieldGeneration != m_field->generation()))
    dropDeviceVolume(); // (volumes over distributed fields are rebuilt on every finalize)
  if (!m_known) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "unknown volume subtype '%s'", subtype.c_str());
    return;
  }
  if (!m_field) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "missing parameter 'value' on transferFunction1D ANARIVolume");
    return;
  }
  if (!m_field->isValid())
    return;
  // discritizeTFData
  const float *color = nullptr, *opacity = nullptr;
  size_t nColor = 0, nOpacity = 0;
  int channels = 4;
  if (m_color) {
    if (m_color->onDevice())
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "tf1D color array must be host memory");
    else if (m_color->elementType == ANARI_FLOAT32_VEC3 || m_color->elementType == ANARI_FLOAT32_VEC4) {
      color = (const float *)m_color->regionData();
      nColor = m_color->regionSize();
      channels = m_color->elementType == ANARI_FLOAT32_VEC3 ? 3 : 4;
    } else
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "unusable tf1D color array type set (%s)",
          typeName(m_color->elementType));
  }
  if (m_opacity && !m_opacity->onDevice() && m_opacity->elementType == ANARI_FLOAT32) {
    opacity = (const float *)m_opacity->regionData();
    nOpacity = m_opacity->regionSize();
  }
  std::vector<float> tf(DVR_TF_SIZE * 4);
  if (dvr_tf_discretize(color, nColor, channels, opacity, nOpacity, m_uniformColor, m_uniformOpacity, m_valueRange,
          tf.data())
      != DVR_OK) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "transfer function rejected: %s", dvr_last_error());
    dropDeviceVolume(); // a rejected edit must not leave the previous table rendering as if it were current
    return;
  }
  m_tf = tf;
  if (m_field->distributed()) { // one device volume per GPU, over that GPU's slab / replica
    const int world = device->gpuCount();
    m_vparts.assign((size_t)world, nullptr);
    for (int r = 0; r < world; ++r) {
      GpuScope scope(device->gpu(r));
      const int rc = dvr_volume_create(m_field->part(r), tf.data(), m_valueRange, m_unitDistance, m_id, device->stream(r),
          &m_vparts[(size_t)r]);
      if (rc != DVR_OK) {
        m_vparts[(size_t)r] = nullptr;
        report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "volume upload to GPU %d failed: %s", device->gpu(r),
            dvr_last_error());
        dropDeviceVolume();
        return;
      }
    }
    if (!m_field->slabbed())
      m_volume = m_vparts[0]; // the display GPU's replica is the whole volume
    m_volumeField = m_field->handle();
    m_volumeFieldGeneration = m_field->generation();
    return;
  }
  CudaDeviceScope scope(device);
  int rc = DVR_ERR_UNSUPPORTED;
  if (m_volume) // same field object, same generation (checked above): refresh in place
    rc = dvr_volume_update(m_volume, tf.data(), m_valueRange, m_unitDistance, m_id, device->stream());
  if (rc == DVR_ERR_UNSUPPORTED) { // no volume yet, or the field's macrocell grid no longer matches the allocation
    dropDeviceVolume();
    rc = dvr_volume_create(m_field->handle(), tf.data(), m_valueRange, m_unitDistance, m_id, device->stream(), &m_volume);
    if (rc != DVR_OK)
      m_volume = nullptr;
    m_volumeField = m_field->handle();
    m_volumeFieldGeneration = m_field->generation();
  }
  if (rc != DVR_OK)
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "volume upload failed: %s", dvr_last_error());
}

DvrVolume *Volume::whole()
{
  if (m_volume || m_vparts.empty() || !m_field)
    return m_volume;
  DvrField *f = m_field->whole();
  if (!f || m_tf.empty())
    return nullptr;
  CudaDeviceScope scope(device);
  if (dvr_volume_create(f, m_tf.data(), m_valueRange, m_unitDistance, m_id, device->stream(), &m_volume) != DVR_OK) {
    m_volume = nullptr;
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "volume upload failed: %s", dvr_last_error());
  }
  return m_volume;
}

// ---------------------------------------------------------------------------------------------------------
// Group / Instance / World
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// Geometry / Material / Surface / Light
// ---------------------------------------------------------------------------------------------------------
Geometry::Geometry(Device *d, const std::string &st) : Object(d, ANARI_GEOMETRY, st)
{
  kind = st == "triangle" ? DVR_GEOMETRY_TRIANGLE : (st == "sphere" ? DVR_GEOMETRY_SPHERE : -1);
}

void Geometry::commitParameters()
{
  auto typed = [this](const char *name, ANARIDataType want) -> Array * {
    Array *a = static_cast<Array *>(getParamObject(name, ANARI_ARRAY1D));
    if (a && a->elementType != want) {
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "'%s' on %s geometry has element type %s, expected %s",
          name, subtype.c_str(), typeName(a->elementType), typeName(want));
      return nullptr;
    }
    if (a && a->onDevice()) {
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
          "'%s' on %s geometry lives in device memory: geometry arrays must be host arrays", name, subtype.c_str());
      return nullptr;
    }
    return a;
  };
  if (kind < 0) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "geometry subtype '%s' is not rendered by this device (triangle and sphere are)", subtype.c_str());
    return;
  }
  const bool tri = kind == DVR_GEOMETRY_TRIANGLE;
  m_vertex.reset(typed("vertex.position", ANARI_FLOAT32_VEC3));
  m_index.reset(typed("primitive.index", tri ? ANARI_UINT32_VEC3 : ANARI_UINT32));
  m_normal.reset(tri ? typed("vertex.normal", ANARI_FLOAT32_VEC3) : nullptr);
  m_radius.reset(tri ? nullptr : typed("vertex.radius", ANARI_FLOAT32));
  m_primId.reset(typed("primitive.id", ANARI_UINT32));
  radius = getParam<float>("radius", ANARI_FLOAT32, 0.01f);
  cullBackfaces = getParam<int32_t>("cullBackfaces", ANARI_BOOL, 0) != 0;
  if (!m_vertex)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "missing required parameter 'vertex.position' on %s geometry",
        subtype.c_str());
  else if (tri && !m_index && m_vertex->regionSize() % 3 != 0) { // Triangle.cpp:64-70
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_ARGUMENT,
        "'vertex.position' on triangle geometry is a non-multiple of 3 without 'primitive.index' present");
    m_vertex.reset();
  }
  if (m_normal && m_vertex && m_normal->regionSize() != m_vertex->regionSize()) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "'vertex.normal' on triangle geometry not the same size as 'vertex.position'");
    m_normal.reset();
  }
  if (m_radius && m_vertex && m_radius->regionSize() != m_vertex->regionSize())
    m_radius.reset();
  const size_t nPrims = m_index ? m_index->regionSize() : (m_vertex ? (tri ? m_vertex->regionSize() / 3 : m_vertex->regionSize()) : 0);
  if (m_primId && m_primId->regionSize() < nPrims)
    m_primId.reset();
}

Material::Material(Device *d, const std::string &st) : Object(d, ANARI_MATERIAL, st) {}

void Material::commitParameters()
{
  color[0] = color[1] = color[2] = 0.8f;
  color[3] = 1.f;
  const bool pbr = subtype == "physicallyBased";
  getParamRaw(pbr ? "baseColor" : "color", ANARI_FLOAT32_VEC4, color, 16);
  getParamRaw(pbr ? "baseColor" : "color", ANARI_FLOAT32_VEC3, color, 12);
  opacity = getParam<float>("opacity", ANARI_FLOAT32, 1.f);
  alphaCutoff = getParam<float>("alphaCutoff", ANARI_FLOAT32, 0.5f);
  const std::string mode = getParamString("alphaMode", "opaque");
  alphaMode = mode == "blend" ? DVR_ALPHA_BLEND : (mode == "mask" ? DVR_ALPHA_MASK : DVR_ALPHA_OPAQUE);
  if (subtype != "matte")
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "material subtype '%s' is shaded as matte by this device", subtype.c_str());
  const Param *c = findParam(pbr ? "baseColor" : "color");
  if (c && (c->type == ANARI_STRING || c->type == ANARI_SAMPLER))
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "material colour from an attribute or sampler is not built: the default colour is used");
}

Surface::Surface(Device *d) : Object(d, ANARI_SURFACE) {}

void Surface::commitParameters()
{
  id = getParam<uint32_t>("id", ANARI_UINT32, ~0u);
  m_geometry.reset(static_cast<Geometry *>(getParamObject("geometry", ANARI_GEOMETRY)));
  m_material.reset(static_cast<Material *>(getParamObject("material", ANARI_MATERIAL)));
  if (!m_material)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "missing 'material' on ANARISurface");
  if (!m_geometry)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "missing 'geometry' on ANARISurface");
}

Light::Light(Device *d, const std::string &st) : Object(d, ANARI_LIGHT, st)
{
  kind = st == "directional" ? DVR_LIGHT_DIRECTIONAL : (st == "point" ? DVR_LIGHT_POINT : -1);
}

void Light::commitParameters()
{
  color[0] = color[1] = color[2] = 1.f;
  getParamRaw("color", ANARI_FLOAT32_VEC3, color, 12);
  if (kind == DVR_LIGHT_DIRECTIONAL) { // Directional.cpp:40-48
    float d[3] = {0.f, 0.f, -1.f};
    getParamRaw("direction", ANARI_FLOAT32_VEC3, d, 12);
    const float inv = 1.f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int i = 0; i < 3; ++i)
      vec[i] = d[i] * inv;
    strength = std::max(getParam<float>("irradiance", ANARI_FLOAT32, 1.f), 0.f);
  } else if (kind == DVR_LIGHT_POINT) { // Point.cpp:40-48
    vec[0] = vec[1] = vec[2] = 0.f;
    getParamRaw("position", ANARI_FLOAT32_VEC3, vec, 12);
    strength = std::max(getParam<float>("intensity", ANARI_FLOAT32, getParam<float>("power", ANARI_FLOAT32, 1.f)), 0.f);
  } else
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "light subtype '%s' is not rendered by this device (directional and point are)", subtype.c_str());
}

Group::Group(Device *d) : Object(d, ANARI_GROUP) {}
void Group::commitParameters()
{
  m_volumes.reset(static_cast<Array *>(getParamObject("volume", ANARI_ARRAY1D)));
  m_surfaces.reset(static_cast<Array *>(getParamObject("surface", ANARI_ARRAY1D)));
  m_lights.reset(static_cast<Array *>(getParamObject("light", ANARI_ARRAY1D)));
}
template <typename T>
static std::vector<T *> objectsOf(const Ref<Array> &a, ANARIDataType type)
{
  std::vector<T *> out;
  if (a)
    for (size_t i = 0; i < a->regionSize(); ++i) {
      Object *o = a->regionObjectAt(i);
      if (o && o->type == type)
        out.push_back(static_cast<T *>(o));
    }
  return out;
}
std::vector<Surface *> Group::surfaces() const { return objectsOf<Surface>(m_surfaces, ANARI_SURFACE); }
std::vector<Light *> Group::lights() const { return objectsOf<Light>(m_lights, ANARI_LIGHT); }
std::vector<Volume *> Group::volumes() const
{
  std::vector<Volume *> out;
  if (m_volumes)
    for (size_t i = 0; i < m_volumes->regionSize(); ++i) {
      Object *o = m_volumes->regionObjectAt(i);
      if (o && o->type == ANARI_VOLUME)
        out.push_back(static_cast<Volume *>(o));
    }
  return out;
}

Instance::Instance(Device *d, const std::string &st) : Object(d, ANARI_INSTANCE, st)
{
  static const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  std::memcpy(objectToWorld, ident, sizeof(ident));
}

void Instance::commitParameters()
{ // scene/Instance.cpp:40-46
  id = getParam<uint32_t>("id", ANARI_UINT32, ~0u);
  static const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  std::memcpy(objectToWorld, ident, sizeof(ident));
  float m4[16];
  if (getParamRaw("transform", ANARI_FLOAT32_MAT4, m4, sizeof(m4))) {
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 3; ++r)
        objectToWorld[c * 3 + r] = m4[c * 4 + r];
  }
  getParamRaw("transform", ANARI_FLOAT32_MAT3x4, objectToWorld, sizeof(objectToWorld));
  m_group.reset(static_cast<Group *>(getParamObject("group", ANARI_GROUP)));
  if (!m_group)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "missing 'group' on ANARIInstance");
}

World::World(Device *d) : Object(d, ANARI_WORLD) {}
World::~World() { dropSurfaceSet(); }
void World::commitParameters()
{
  m_zeroVolumes.reset(static_cast<Array *>(getParamObject("volume", ANARI_ARRAY1D)));
  m_instances.reset(static_cast<Array *>(getParamObject("instance", ANARI_ARRAY1D)));
  m_zeroSurfaces.reset(static_cast<Array *>(getParamObject("surface", ANARI_ARRAY1D)));
  m_zeroLights.reset(static_cast<Array *>(getParamObject("light", ANARI_ARRAY1D)));
}

void World::dropSurfaceSet()
{
  if (!m_surfaceSet)
    return;
  GpuScope scope(m_surfaceSetGpu);
  cudaDeviceSynchronize();
  dvr_surfaces_destroy(m_surfaceSet);
  m_surfaceSet = nullptr;
  m_surfaceStamp = 0;
}

std::vector<World::FlatSurface> World::flattenSurfaces(bool warn) const
{
  std::vector<FlatSurface> out;
  static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  auto push = [&](Surface *s, const float *o2w, uint32_t instId) {
    if (!s->isValid()) {
      if (warn)
        report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "skipping invalid surface in world");
      return;
    }
    FlatSurface fs;
    fs.surface = s;
    std::memcpy(fs.objectToWorld, o2w, sizeof(fs.objectToWorld));
    fs.instId = instId;
    out.push_back(fs);
  };
  for (Surface *s : objectsOf<Surface>(m_zeroSurfaces, ANARI_SURFACE))
    push(s, ident, ~0u);
  for (Instance *in : objectsOf<Instance>(m_instances, ANARI_INSTANCE)) {
    if (!in->isValid())
      continue;
    float rm[12]; // column-major 4x3 -> row-major 3x4
    for (int r = 0; r < 3; ++r) {
      rm[r * 4 + 0] = in->objectToWorld[0 * 3 + r];
      rm[r * 4 + 1] = in->objectToWorld[1 * 3 + r];
      rm[r * 4 + 2] = in->objectToWorld[2 * 3 + r];
      rm[r * 4 + 3] = in->objectToWorld[9 + r];
    }
    for (Surface *s : in->group()->surfaces())
      push(s, rm, in->id);
  }
  return out;
}

std::vector<DvrLight> World::flattenLights() const
{
  std::vector<DvrLight> out;
  auto push = [&](const Light *l, const float *cm /*column-major 4x3 or null*/) {
    if (!l->isValid())
      return;
    DvrLight d;
    d.type = l->kind;
    std::memcpy(d.color, l->color, sizeof(d.color));
    d.strength = l->strength;
    for (int r = 0; r < 3; ++r) {
      if (!cm)
        d.vec[r] = l->vec[r];
      else // xfmVec for directions, xfmPoint for positions (gpu/gpu_math.h:314-322)
        d.vec[r] = cm[0 * 3 + r] * l->vec[0] + cm[1 * 3 + r] * l->vec[1] + cm[2 * 3 + r] * l->vec[2]
            + (l->kind == DVR_LIGHT_POINT ? cm[9 + r] : 0.f);
    }
    out.push_back(d);
  };
  for (Light *l : objectsOf<Light>(m_zeroLights, ANARI_LIGHT))
    push(l, nullptr);
  for (Instance *in : objectsOf<Instance>(m_instances, ANARI_INSTANCE))
    if (in->isValid())
      for (Light *l : in->group()->lights())
        push(l, in->objectToWorld);
  return out;
}

DvrSurfaces *World::surfaceSet(bool warn)
{
  const std::vector<FlatSurface> flat = flattenSurfaces(warn);
  if (flat.empty()) {
    dropSurfaceSet();
    return nullptr;
  }
  // fingerprint: the finalisation stamps of everything the set is built from (+ identities and transforms)
  uint64_t stamp = 1469598103934665603ull;
  auto mix = [&stamp](uint64_t v) { stamp = (stamp ^ v) * 1099511628211ull; };
  auto mixArray = [&](const Array *a) { mix(a ? a->lastFinalized + 1 : 0); mix((uint64_t)(uintptr_t)a); };
  for (const FlatSurface &fs : flat) {
    const Geometry *g = fs.surface->geometry();
    mix((uint64_t)(uintptr_t)fs.surface);
    mix(fs.surface->lastFinalized);
    mix(g->lastFinalized);
    mix(fs.surface->material()->lastFinalized);
    mixArray(g->vertex());
    mixArray(g->index());
    mixArray(g->normal());
    mixArray(g->vertexRadius());
    mixArray(g->primitiveId());
    for (int i = 0; i < 12; ++i) {
      uint32_t bits;
      std::memcpy(&bits, &fs.objectToWorld[i], 4);
      mix(bits);
    }
    mix(fs.instId);
  }
  if (m_surfaceSet && stamp == m_surfaceStamp && m_surfaceSetGpu == device->cudaDevice())
    return m_surfaceSet;
  dropSurfaceSet();
  std::vector<DvrSurfaceDesc> descs(flat.size());
  for (size_t i = 0; i < flat.size(); ++i) {
    const Geometry *g = flat[i].surface->geometry();
    const Material *m = flat[i].surface->material();
    DvrSurfaceDesc &d = descs[i];
    std::memset(&d, 0, sizeof(d));
    d.geometryType = g->kind;
    d.nVertices = (uint32_t)g->vertex()->regionSize();
    d.vertexPosition = (const float *)g->vertex()->regionData();
    d.index = g->index() ? (const uint32_t *)g->index()->regionData() : nullptr;
    d.nPrimitives = g->index() ? (uint32_t)g->index()->regionSize()
                               : (g->kind == DVR_GEOMETRY_TRIANGLE ? d.nVertices / 3u : d.nVertices);
    d.vertexNormal = g->normal() ? (const float *)g->normal()->regionData() : nullptr;
    d.vertexRadius = g->vertexRadius() ? (const float *)g->vertexRadius()->regionData() : nullptr;
    d.radius = g->radius;
    d.primitiveId = g->primitiveId() ? (const uint32_t *)g->primitiveId()->regionData() : nullptr;
    d.cullBackfaces = g->cullBackfaces ? 1 : 0;
    std::memcpy(d.color, m->color, sizeof(d.color));
    d.opacity = m->opacity;
    d.alphaMode = m->alphaMode;
    d.alphaCutoff = m->alphaCutoff;
    d.surfaceId = flat[i].surface->id;
    d.instanceId = flat[i].instId;
    std::memcpy(d.objectToWorld, flat[i].objectToWorld, sizeof(d.objectToWorld));
  }
  const int rc = dvr_surfaces_create(descs.data(), (uint32_t)descs.size(), device->stream(), &m_surfaceSet);
  if (rc != DVR_OK) {
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "dvr_surfaces_create failed: %s", dvr_last_error());
    m_surfaceSet = nullptr;
    return nullptr;
  }
  m_surfaceStamp = stamp;
  m_surfaceSetGpu = device->cudaDevice();
  return m_surfaceSet;
}

// inverse of an affine transform given as column-major 4x3 -> row-major 3x4
static bool invertAffine(const float m[12], float out[12])
{
  const double a = m[0], b = m[3], c = m[6], d = m[1], e = m[4], f = m[7], g = m[2], h = m[5], i = m[8];
  const double tx = m[9], ty = m[10], tz = m[11];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  if (det == 0.0)
    return false;
  const double r[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det, (f * g - d * i) / det,
      (a * i - c * g) / det, (c * d - a * f) / det, (d * h - e * g) / det, (b * g - a * h) / det,
      (a * e - b * d) / det};
  for (int row = 0; row < 3; ++row) {
    out[row * 4 + 0] = (float)r[row * 3 + 0];
    out[row * 4 + 1] = (float)r[row * 3 + 1];
    out[row * 4 + 2] = (float)r[row * 3 + 2];
    out[row * 4 + 3] = (float)(-(r[row * 3 + 0] * tx + r[row * 3 + 1] * ty + r[row * 3 + 2] * tz));
  }
  return true;
}

std::vector<FlatInstance> World::flatten(bool warn) const
{
  std::vector<FlatInstance> out;
  static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  auto push = [&](Volume *v, const float *w2o, uint32_t instId) {
    if (!v->isValid()) {
      if (warn)
        report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "skipping invalid volume in world");
      return;
    }
    FlatInstance fi;
    fi.volume = v;
    std::memcpy(fi.worldToObject, w2o, sizeof(fi.worldToObject));
    fi.instId = instId;
    out.push_back(fi);
  };
  if (m_zeroVolumes) // the world's own volumes live in an identity "zero instance" (World.cpp:82-96,117-135)
    for (size_t i = 0; i < m_zeroVolumes->regionSize(); ++i) {
      Object *o = m_zeroVolumes->regionObjectAt(i);
      if (o && o->type == ANARI_VOLUME)
        push(static_cast<Volume *>(o), ident, ~0u);
    }
  if (m_instances)
    for (size_t i = 0; i < m_instances->regionSize(); ++i) {
      Object *o = m_instances->regionObjectAt(i);
      if (!o || o->type != ANARI_INSTANCE)
        continue;
      Instance *in = static_cast<Instance *>(o);
      if (!in->isValid()) {
        if (warn)
          report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "skipping invalid instance in world");
        continue;
      }
      float w2o[12];
      if (!invertAffine(in->objectToWorld, w2o)) {
        if (warn)
          report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "singular instance transform");
        continue;
      }
      for (Volume *v : in->group()->volumes())
        push(v, w2o, in->id);
    }
  return out;
}

void World::bounds(float lo[3], float hi[3]) const
{
  const float big = std::numeric_limits<float>::max();
  lo[0] = lo[1] = lo[2] = big;
  hi[0] = hi[1] = hi[2] = -big;
  for (const FlatInstance &fi : flatten(false)) {
    float flo[3], fhi[3];
    fi.volume->field()->bounds(flo, fhi);
    // object-space corners -> world: invert the stored world->object (row-major 3x4)
    float o2w[12];
    float cm[12];
    for (int r = 0; r < 3; ++r) {
      cm[0 * 3 + r] = fi.worldToObject[r * 4 + 0];
      cm[1 * 3 + r] = fi.worldToObject[r * 4 + 1];
      cm[2 * 3 + r] = fi.worldToObject[r * 4 + 2];
      cm[9 + r] = fi.worldToObject[r * 4 + 3];
    }
    if (!invertAffine(cm, o2w))
      continue;
    for (int k = 0; k < 8; ++k) {
      const float p[3] = {k & 1 ? fhi[0] : flo[0], k & 2 ? fhi[1] : flo[1], k & 4 ? fhi[2] : flo[2]};
      for (int r = 0; r < 3; ++r) {
        const float w = o2w[r * 4] * p[0] + o2w[r * 4 + 1] * p[1] + o2w[r * 4 + 2] * p[2] + o2w[r * 4 + 3];
        lo[r] = std::min(lo[r], w);
        hi[r] = std::max(hi[r], w);
      }
    }
  }
}

// object-space box of a geometry's primitives (triangle corners / spheres)
static bool geometryBounds(const Geometry *g, float lo[3], float hi[3])
{
  if (!g || !g->isValid())
    return false;
  const float *v = (const float *)g->vertex()->regionData();
  const size_t n = g->vertex()->regionSize();
  const float *rad = g->vertexRadius() ? (const float *)g->vertexRadius()->regionData() : nullptr;
  for (int a = 0; a < 3; ++a) {
    lo[a] = std::numeric_limits<float>::max();
    hi[a] = -std::numeric_limits<float>::max();
  }
  for (size_t i = 0; i < n; ++i) {
    const float r = g->kind == DVR_GEOMETRY_SPHERE ? std::fabs(rad ? rad[i] : g->radius) : 0.f;
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], v[3 * i + a] - r);
      hi[a] = std::max(hi[a], v[3 * i + a] + r);
    }
  }
  return n > 0;
}

bool World::getProperty(const std::string &name, ANARIDataType t, void *mem, uint64_t size, uint32_t mask)
{
  if (name == "bounds" && t == ANARI_FLOAT32_BOX3 && size >= 24) { // World.cpp:100-117
    if (mask & ANARI_WAIT)
      device->flushCommits();
    float b[6];
    bounds(b, b + 3);
    for (const FlatSurface &fs : flattenSurfaces(false)) { // surfaces extend the box
      float lo[3], hi[3];
      if (!geometryBounds(fs.surface->geometry(), lo, hi))
        continue;
      for (int k = 0; k < 8; ++k) {
        const float p[3] = {k & 1 ? hi[0] : lo[0], k & 2 ? hi[1] : lo[1], k & 4 ? hi[2] : lo[2]};
        for (int r = 0; r < 3; ++r) {
          const float *m = fs.objectToWorld + 4 * r;
          const float w = m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + m[3];
          b[r] = std::min(b[r], w);
          b[3 + r] = std::max(b[3 + r], w);
        }
      }
    }
    std::memcpy(mem, b, 24);
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------
// Renderer (renderer/Renderer.cpp:152-170, Raycast.cpp:45-50)
// ---------------------------------------------------------------------------------------------------------
Renderer::Renderer(Device *d, const std::string &st) : Object(d, ANARI_RENDERER, st)
{
  std::string s = st;
  if (const char *ov = getenv("VISRTX_OVERRIDE_RENDERER")) // Renderer.cpp:263-266
    s = ov;
  subtype = s;
  m_known = s == "raycast" || s == "default" || s == "directLight" || s == "ao" || s == "dpt"
      || s == "diffuse_pathtracer" || s == "test";
  d->enqueueCommit(this);
}

void Renderer::commitParameters()
{
  background[0] = background[1] = background[2] = 0.f;
  background[3] = 1.f;
  getParamRaw("background", ANARI_FLOAT32_VEC4, background, 16);
  { // "background" as an Array2D (Renderer.cpp:154): sampled per pixel-sample; the texture is (re)built in finalize
    Array *img = static_cast<Array *>(getParamObject("background", ANARI_ARRAY2D));
    if (m_bgArray.ptr != img) {
      if (m_bgArray)
        m_bgArray->removeObserver(this);
      m_bgArray.reset(img);
      if (img)
        img->addObserver(this);
    }
  }
  spp = getParam<int>("pixelSamples", ANARI_INT32, 1);
  checkerboard = getParam<int32_t>("checkerboarding", ANARI_BOOL, 0) != 0;
  sampleLimit = getParam<int>("sampleLimit", ANARI_INT32, 128);
  volumeSamplingRate = std::min(std::max(getParam<float>("volumeSamplingRate", ANARI_FLOAT32, 0.125f), 1e-3f), 10.f);
  { // extension; image-neutral.  Unset: the device decides per volume (DVR_SKIP_AUTO)
    const int32_t ms = getParam<int32_t>("macrocellSkipping", ANARI_BOOL, -1);
    macrocellSkipping = ms < 0 ? DVR_SKIP_AUTO : (ms ? DVR_SKIP_ON : DVR_SKIP_OFF);
  }
  tileRank = (uint32_t)std::max(getParam<int>("sortFirstRank", ANARI_INT32, 0), 0);
  tileRanks = (uint32_t)std::max(getParam<int>("sortFirstRanks", ANARI_INT32, 1), 1);
  if (checkerboard)
    spp = 1;
  integrator = DVR_INTEGRATOR_DEFAULT;
  if (subtype == "raycast") {
    integrator = DVR_INTEGRATOR_RAYCAST;
    sampleLimit = 1; // single-shot renderer
  }
  if (subtype == "test") // Test_ptx.cu: colour = primary ray direction
    integrator = DVR_INTEGRATOR_TEST;
  if (subtype == "dpt" || subtype == "diffuse_pathtracer") { // DiffusePathTracer.cpp:41-53
    integrator = DVR_INTEGRATOR_DPT;
    maxDepth = std::min(std::max(getParam<int>("maxDepth", ANARI_INT32, 5), 1), 256);
    // extension: walk the grid content the reference itself builds (Q7/Q8) instead of the conservative one
    dptReferenceGrid = getParam<int32_t>("dptReferenceGrid", ANARI_BOOL, 0) != 0;
  }
  // Renderer.cpp:159-161; the dpt renderer's default ambient radiance is 1 (DiffusePathTracer.cpp:41)
  ambientRadiance = getParam<float>("ambientRadiance", ANARI_FLOAT32, integrator == DVR_INTEGRATOR_DPT ? 1.f : 0.f);
  occlusionDistance = getParam<float>("ambientOcclusionDistance", ANARI_FLOAT32, 1e20f);
  // surface shading of mixed scenes (Renderer.cpp:158,165, DirectLight.cpp:49-53)
  ambientColor[0] = ambientColor[1] = ambientColor[2] = 1.f;
  getParamRaw("ambientColor", ANARI_FLOAT32_VEC3, ambientColor, 12);
  ambientSamples = std::min(std::max(getParam<int>("ambientSamples", ANARI_INT32, 1), 0), 256);
  cullTriangleBackfaces = getParam<int32_t>("cullTriangleBackfaces", ANARI_BOOL, 0) != 0;
  if (!m_known)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "unknown renderer subtype '%s'", subtype.c_str());
}

Renderer::~Renderer()
{
  if (m_bgArray)
    m_bgArray->removeObserver(this);
  dropBackgroundImage();
}

void Renderer::dropBackgroundImage()
{
  if (!m_bgImage)
    return;
  CudaDeviceScope scope(device);
  cudaStreamSynchronize((cudaStream_t)device->stream());
  dvr_image_destroy(m_bgImage);
  m_bgImage = nullptr;
}

// Renderer::finalize, renderer/Renderer.cpp:172-179: Array2D::acquireCUDAArrayUint8 + a clamp / linear texture.
// Element types: what numANARIChannels / isFloat / isFixed8|16|32 / isSrgb8 accept (utility/AnariTypeHelpers.h).
void Renderer::finalize()
{
  dropBackgroundImage();
  if (!m_bgArray)
    return;
  const ANARIDataType t = m_bgArray->elementType;
  int comp = -1, channels = 0;
  if (t >= ANARI_FLOAT32 && t <= ANARI_FLOAT32_VEC4) {
    comp = DVR_IMAGE_FLOAT32;
    channels = t - ANARI_FLOAT32 + 1;
  } else if (t >= ANARI_UFIXED8 && t <= ANARI_UFIXED8_VEC4) {
    comp = DVR_IMAGE_UFIXED8;
    channels = t - ANARI_UFIXED8 + 1;
  } else if (t >= ANARI_UFIXED16 && t <= ANARI_UFIXED16_VEC4) {
    comp = DVR_IMAGE_UFIXED16;
    channels = t - ANARI_UFIXED16 + 1;
  } else if (t >= ANARI_UFIXED32 && t <= ANARI_UFIXED32_VEC4) {
    comp = DVR_IMAGE_UFIXED32;
    channels = t - ANARI_UFIXED32 + 1;
  } else if (t >= ANARI_UFIXED8_R_SRGB && t <= ANARI_UFIXED8_RGBA_SRGB) {
    comp = DVR_IMAGE_SRGB8;
    channels = t - ANARI_UFIXED8_R_SRGB + 1;
  }
  if (comp < 0 || m_bgArray->dims[0] == 0 || m_bgArray->dims[1] == 0) {
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT,
        "unusable background image element type (%d); using the constant colour", (int)t);
    return;
  }
  if (!device->initDevice())
    return;
  CudaDeviceScope scope(device);
  const void *pixels = m_bgArray->data();
  std::vector<uint8_t> staged;
  if (m_bgArray->onDevice()) { // the staging pass reads host memory (Array::data() of the reference is a host view)
    staged.resize(m_bgArray->totalBytes());
    cudaMemcpy(staged.data(), pixels, staged.size(), cudaMemcpyDeviceToHost);
    pixels = staged.data();
  }
  if (dvr_image_create(pixels, comp, channels, (uint32_t)m_bgArray->dims[0], (uint32_t)m_bgArray->dims[1],
          device->stream(), &m_bgImage)
      != DVR_OK) {
    m_bgImage = nullptr;
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "background image upload failed: %s", dvr_last_error());
  }
}

} // namespace b200
