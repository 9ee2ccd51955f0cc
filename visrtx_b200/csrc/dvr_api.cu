// dvr_api.cu — the extern "C" launch layer declared in include/dvr_b200.h.
// Host-side object state (3-D array + texture objects, macrocell buffers, TF tables) and the
// parameter translation from the POD structs of the C-ABI to the kernel parameter blocks.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "dvr_internal.h"
#include "../../include/dvr_nvdb_validate.h"

namespace dvr {

static thread_local std::string g_lastError;
static std::atomic<unsigned long long> g_launches{0};

void setError(const std::string &msg) { g_lastError = msg; }

int cudaFail(cudaError_t e, const char *what)
{
  g_lastError = std::string(what) + ": " + cudaGetErrorString(e);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
    return DVR_ERR_NO_DEVICE;
  if (e == cudaErrorMemoryAllocation)
    return DVR_ERR_OUT_OF_MEMORY;
  return DVR_ERR_CUDA;
}

void countLaunch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// per-device scheduler counters: a ring of {nextTile, warpsDone, nextRowItem, pad} slots, always zero between launches
struct DeviceScratch
{
  unsigned int *sched = nullptr;
  unsigned next = 0;
  int sms = 0;
  InstanceDev *extInstances = nullptr;
  size_t extCapacity = 0;
  // auxiliary stream of the background sweep + fork / join events (created on first use)
  cudaStream_t sweepStream = nullptr;
  cudaEvent_t sweepFork = nullptr, sweepJoin = nullptr;
  // stream-ordered scratch (macrocell build intermediates, visibility bitmaps ...): a private pool that keeps its
  // pages across synchronisations.  The default pool hands freed memory back to the driver at every sync, which
  // costs milliseconds per field refresh.
  cudaMemPool_t pool = nullptr;
};
static std::mutex g_scratchMutex;
static DeviceScratch g_scratch[64];
constexpr unsigned kSchedSlots = 256;

static DeviceScratch *scratch()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return nullptr;
  std::lock_guard<std::mutex> lock(g_scratchMutex);
  DeviceScratch &s = g_scratch[dev];
  if (!s.sched) {
    if (cudaMalloc(&s.sched, kSchedSlots * 4 * sizeof(unsigned int)) != cudaSuccess)
      return nullptr;
    cudaMemset(s.sched, 0, kSchedSlots * 4 * sizeof(unsigned int));
    cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev);
    // DRAM fetch granularity of L2 misses (32 / 64 / 128 B).  The march reads the whole block-linear volume through
    // scattered 32-byte sectors; a wider fill turns neighbouring sector misses into L2 hits.  A/B knob, measured in
    // profiles/r02_l2_fetch_granularity.md.
    if (const char *g = std::getenv("DVR_B200_L2_FETCH")) {
      const int bytes = std::atoi(g);
      if (bytes == 32 || bytes == 64 || bytes == 128)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
    }
  }
  return &s;
}

cudaError_t scratchAllocAsync(void **p, size_t bytes, cudaStream_t stream)
{
  DeviceScratch *s = scratch();
  if (s) {
    std::lock_guard<std::mutex> lock(g_scratchMutex);
    if (!s->pool) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaMemPoolProps props;
      std::memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if (cudaMemPoolCreate(&s->pool, &props) == cudaSuccess) {
        uint64_t keep = 1ull << 30;
        cudaMemPoolSetAttribute(s->pool, cudaMemPoolAttrReleaseThreshold, &keep);
      } else {
        cudaGetLastError();
        s->pool = nullptr;
      }
    }
  }
  if (s && s->pool)
    return cudaMallocFromPoolAsync(p, bytes, s->pool, stream);
  return cudaMallocAsync(p, bytes, stream);
}

unsigned int *acquireSchedSlot()
{
  DeviceScratch *s = scratch();
  if (!s)
    return nullptr;
  std::lock_guard<std::mutex> lock(g_scratchMutex);
  unsigned int *p = s->sched + 4 * (s->next % kSchedSlots);
  s->next++;
  return p;
}

// Runs the background sweep of one frame on the device's auxiliary stream, ordered after everything already
// enqueued on `s`; the caller enqueues the frame kernel on `s` and then calls joinBackgroundSweep.
// The auxiliary stream first waits for everything already enqueued on `s` (fork), the sweep kernel is then enqueued
// on it before or after the frame kernel's own launch on `s`, and `s` finally waits for the sweep (join).
static int forkBackgroundSweep(cudaStream_t s)
{
  DeviceScratch *sc = scratch();
  if (!sc)
    return DVR_ERR_CUDA;
  {
    std::lock_guard<std::mutex> lock(g_scratchMutex);
    if (!sc->sweepStream) {
      DVR_CUDA(cudaStreamCreateWithFlags(&sc->sweepStream, cudaStreamNonBlocking));
      DVR_CUDA(cudaEventCreateWithFlags(&sc->sweepFork, cudaEventDisableTiming));
      DVR_CUDA(cudaEventCreateWithFlags(&sc->sweepJoin, cudaEventDisableTiming));
    }
  }
  DVR_CUDA(cudaEventRecord(sc->sweepFork, s));
  DVR_CUDA(cudaStreamWaitEvent(sc->sweepStream, sc->sweepFork, 0));
  return DVR_OK;
}

static int enqueueBackgroundSweep(const FrameLaunch &L)
{
  DeviceScratch *sc = scratch();
  if (!sc || !sc->sweepStream)
    return DVR_ERR_CUDA;
  const int rc = launchBackgroundSweep(L, sc->sweepStream);
  DVR_CUDA(cudaEventRecord(sc->sweepJoin, sc->sweepStream));
  return rc;
}

static int joinBackgroundSweep(cudaStream_t s)
{
  DeviceScratch *sc = scratch();
  if (!sc || !sc->sweepJoin)
    return DVR_ERR_CUDA;
  DVR_CUDA(cudaStreamWaitEvent(s, sc->sweepJoin, 0));
  return DVR_OK;
}

int smCount()
{
  DeviceScratch *s = scratch();
  return s && s->sms > 0 ? s->sms : 148;
}

} // namespace dvr

using namespace dvr;

// ------------------------------------------------------------------------------------------------
// opaque objects
// ------------------------------------------------------------------------------------------------
struct DvrField
{
  cudaArray_t array = nullptr;
  cudaTextureObject_t tex = 0;      // filtering view (clamp, normalised)
  cudaTextureObject_t pointTex = 0; // point view, unnormalised (macrocell build)
  cudaSurfaceObject_t surf = 0;     // whole f32 fields: store view, so a refresh from device memory is one fused pass
  FieldDev dev{};
  float2 *ranges = nullptr;
  size_t nCells = 0;
  size_t voxelBytes = 0;
  size_t texElementSize = 0;
  void *nvdbBlob = nullptr; // NanoVDB fields: device copy of the serialized grid
  int2 *brickTable = nullptr; // NanoVDB fields: apron bricks (dvr_nvdb_bricks.cu), absent when over budget
  float *bricks = nullptr;
  size_t brickBytes = 0;
  int device = 0;
  int dataType = -1; // structured fields: the DvrDataType handed to create
  int ownLimitBegin = 0, ownLimitEnd = 0; // slab fields: the ownership range given at creation (dvr_field_set_owned_slices)
};

struct DvrVolume
{
  const DvrField *field = nullptr;
  float4 *tf = nullptr;
  float *maxOpacities = nullptr;
  float *maxOpacitiesCoarse = nullptr;
  int3 coarseDims{0, 0, 0};
  float emptyFraction = 0.f;            // share of macrocells with majorant 0 under the current TF
  float *maxOpacitiesCoarse2 = nullptr; // 256^3-voxel blocks
  int3 coarse2Dims{0, 0, 0};
  float *ddaMaxOpacities = nullptr; // delta-tracking grid (built on first use by the dpt integrator)
  float2 *ddaRanges = nullptr;
  bool ddaValid = false;
  bool ddaReference = false; // content of the current grid: the reference's build (Q7/Q8) or the conservative one
  float vrLo = 0.f, vrHi = 1.f, oneOverUnitDistance = 1.f;
  uint32_t id = ~0u;
  // the macrocell grid the majorant buffers were allocated for: an update against a field of another shape is refused
  size_t allocCells = 0;
  int3 allocGridDims{0, 0, 0};
};

struct DvrImage
{
  cudaArray_t array = nullptr;
  cudaTextureObject_t tex = 0;
  uint32_t width = 0, height = 0;
  int channels = 0;
};

static inline float3 v3(const float *p) { return make_float3(p[0], p[1], p[2]); }
static float3 hnormalize(float3 v)
{
  const float d = v.x * v.x + v.y * v.y + v.z * v.z;
  const float inv = 1.0f / std::sqrt(d);
  return make_float3(v.x * inv, v.y * inv, v.z * inv);
}
static float3 hcross(float3 a, float3 b)
{
  return make_float3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static float3 hscale(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
static float3 hsub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static void st3(float *d, float3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

// DVR_SKIP_AUTO: use the skipping kernel when at least this share of the macrocells is empty (on fully dense
// data the test per lattice point costs ~6 %: C2 1068 vs 1007 frames/s)
static constexpr float kSkipAutoThreshold = 1.f / 64.f;

extern "C" {

const char *dvr_last_error(void) { return g_lastError.c_str(); }

int dvr_version(int *major, int *minor)
{
  if (major) *major = DVR_B200_VERSION_MAJOR;
  if (minor) *minor = DVR_B200_VERSION_MINOR;
  return DVR_OK;
}

int dvr_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int dvr_set_device(int cudaDevice)
{
  DVR_CUDA(cudaSetDevice(cudaDevice));
  return DVR_OK;
}

int dvr_device_info(char *name, size_t nameLen, int *sms, size_t *totalMem)
{
  int dev = 0;
  DVR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DVR_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name && nameLen) {
    std::strncpy(name, prop.name, nameLen - 1);
    name[nameLen - 1] = 0;
  }
  if (sms) *sms = prop.multiProcessorCount;
  if (totalMem) *totalMem = prop.totalGlobalMem;
  return DVR_OK;
}

unsigned long long dvr_launch_count(void) { return g_launches.load(); }

// ---- cameras -------------------------------------------------------------------------------------

static void cameraBase(const float pos[3], const float dir[3], const float up[3], const float region[4],
    DvrCamera *out)
{
  std::memset(out, 0, sizeof(*out));
  out->region[0] = region ? region[0] : 0.f;
  out->region[1] = region ? region[1] : 0.f;
  out->region[2] = region ? region[2] : 1.f;
  out->region[3] = region ? region[3] : 1.f;
  st3(out->pos, v3(pos));
  st3(out->dir, hnormalize(v3(dir)));
  st3(out->up, hnormalize(v3(up)));
}

int dvr_camera_perspective(const float pos[3], const float dir[3], const float up[3], float fovy, float aspect,
    float focusDistance, float apertureRadius, const float region[4], DvrCamera *out)
{
  if (!pos || !dir || !up || !out) {
    setError("dvr_camera_perspective: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  cameraBase(pos, dir, up, region, out);
  out->type = DVR_CAMERA_PERSPECTIVE;
  const float3 d = v3(out->dir), u = v3(out->up);
  const float planeY = 2.f * tanf(0.5f * fovy);
  const float planeX = planeY * aspect;
  float3 du = hscale(hnormalize(hcross(d, u)), planeX);
  float3 dv = hscale(hnormalize(hcross(du, d)), planeY);
  float3 d00 = hsub(hsub(d, hscale(du, .5f)), hscale(dv, .5f));
  const float scaledAperture = apertureRadius / (planeX * focusDistance);
  if (scaledAperture > 0.f) {
    du = hscale(du, focusDistance);
    dv = hscale(dv, focusDistance);
    d00 = hscale(d00, focusDistance);
  }
  st3(out->du, du);
  st3(out->dv, dv);
  st3(out->p00, d00);
  out->scaledAperture = scaledAperture;
  out->aspect = aspect;
  return DVR_OK;
}

int dvr_camera_orthographic(const float pos[3], const float dir[3], const float up[3], float height, float aspect,
    const float region[4], DvrCamera *out)
{
  if (!pos || !dir || !up || !out) {
    setError("dvr_camera_orthographic: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  cameraBase(pos, dir, up, region, out);
  out->type = DVR_CAMERA_ORTHOGRAPHIC;
  const float3 d = v3(out->dir), u = v3(out->up), p = v3(out->pos);
  const float3 du = hscale(hnormalize(hcross(d, u)), height * aspect);
  const float3 dv = hscale(hnormalize(hcross(du, d)), height);
  const float3 p00 = hsub(hsub(p, hscale(du, .5f)), hscale(dv, .5f));
  st3(out->du, du);
  st3(out->dv, dv);
  st3(out->p00, p00);
  out->aspect = aspect;
  return DVR_OK;
}

// ---- transfer function discretisation ---------------------------------------------------------------

namespace {
struct Range1
{
  float lo, hi;
};
inline float rsize(Range1 r) { return r.hi - r.lo; }
inline float rposition(float v, Range1 r)
{
  v = std::fmax(r.lo, std::fmin(v, r.hi));
  return (v - r.lo) * (1.f / rsize(r));
}
inline bool rcontains(float v, Range1 r) { return v >= r.lo && v <= r.hi; }

// colorMapHelpers.h:43-59: positions by repeated addition, first/last pinned, mapped into the range
std::vector<float> linearPositions(size_t n, Range1 range)
{
  std::vector<float> p(n);
  p.front() = 0.f;
  p.back() = 1.f;
  const float w = 1.f / (n - 1);
  for (int i = 1; i < (int)p.size() - 1; i++)
    p[i] = p[i - 1] + w;
  for (auto &v : p)
    v = v * rsize(range) + range.lo;
  return p;
}

// colorMapHelpers.h:61-72 for `nch` interleaved channels
void interpolated(const float *values, int nch, const std::vector<float> &positions, Range1 range, float pos,
    float *out)
{
  const size_t n = positions.size();
  for (size_t i = 0; i + 1 < n; i++) {
    const Range1 r{rposition(positions[i], range), rposition(positions[i + 1], range)};
    if (rcontains(pos, r)) {
      const float a = rposition(pos, r);
      for (int c = 0; c < nch; ++c)
        out[c] = values[i * nch + c] * (1.f - a) + values[(i + 1) * nch + c] * a;
      return;
    }
  }
  const float *src = pos <= rposition(positions[0], range) ? values : values + (n - 1) * nch;
  for (int c = 0; c < nch; ++c)
    out[c] = src[c];
}
} // namespace

int dvr_tf_discretize(const float *color, size_t nColor, int colorChannels, const float *opacity, size_t nOpacity,
    const float uniformColor[4], float uniformOpacity, const float valueRange[2], float *outRgba)
{
  if (!outRgba || !valueRange || !uniformColor) {
    setError("dvr_tf_discretize: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (color && (colorChannels != 3 && colorChannels != 4)) {
    setError("dvr_tf_discretize: colorChannels must be 3 or 4");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if ((color && nColor < 2) || (opacity && nOpacity < 2)) {
    setError("dvr_tf_discretize: arrays need at least two entries");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  const Range1 range{valueRange[0], valueRange[1]};
  std::vector<float> cpos, opos;
  if (color)
    cpos = linearPositions(nColor, range);
  if (opacity)
    opos = linearPositions(nOpacity, range);
  for (size_t i = 0; i < DVR_TF_SIZE; i++) {
    const float p = float(i) / (DVR_TF_SIZE - 1);
    float c[4] = {uniformColor[0], uniformColor[1], uniformColor[2], uniformColor[3]};
    if (color) {
      if (colorChannels == 3) {
        interpolated(color, 3, cpos, range, p, c);
        c[3] = 1.f;
      } else
        interpolated(color, 4, cpos, range, p, c);
    }
    float o = uniformOpacity;
    if (opacity)
      interpolated(opacity, 1, opos, range, p, &o);
    outRgba[4 * i + 0] = c[0];
    outRgba[4 * i + 1] = c[1];
    outRgba[4 * i + 2] = c[2];
    outRgba[4 * i + 3] = c[3] * o;
  }
  return DVR_OK;
}

// ---- fields ------------------------------------------------------------------------------------------

static size_t elementSize(int t)
{
  switch (t) {
  case DVR_FLOAT32: return 4;
  case DVR_UFIXED8:
  case DVR_FIXED8: return 1;
  case DVR_UFIXED16:
  case DVR_FIXED16:
  case DVR_FLOAT16: return 2;
  case DVR_FLOAT64: return 8;
  default: return 0;
  }
}

// origin / spacing derived members of a structured field, StructuredRegularField.cpp:166-189
static void setFieldGeometry(FieldDev &d, const uint32_t gdims[3], const float origin[3], const float spacing[3])
{
  d.origin = v3(origin);
  d.spacing = v3(spacing);
  d.dims = make_int3((int)gdims[0], (int)gdims[1], (int)gdims[2]);
  // 1 / (spacing * dims), StructuredRegularField.cpp:188-189
  d.invSpacing = make_float3(1.f / (spacing[0] * (float)gdims[0]), 1.f / (spacing[1] * (float)gdims[1]),
      1.f / (spacing[2] * (float)gdims[2]));
  d.boundsLo = d.origin;
  d.boundsHi = make_float3(origin[0] + ((float)gdims[0] - 1.f) * spacing[0],
      origin[1] + ((float)gdims[1] - 1.f) * spacing[1], origin[2] + ((float)gdims[2] - 1.f) * spacing[2]);
  d.stepSize = std::fmin(std::fmin(spacing[0] / 2.f, spacing[1] / 2.f), spacing[2] / 2.f);
}

static int createFieldImpl(const void *data, int dataIsDevice, int dataType, const uint32_t gdims[3],
    uint32_t zBegin, uint32_t zEnd, const float origin[3], const float spacing[3], int filter, void *stream,
    DvrField **out)
{
  if (!gdims || !origin || !spacing || !out) {
    setError("dvr_field_create: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  const bool deferredUpload = data == nullptr; // slab API only: slices arrive through dvr_field_upload_slices
  if (deferredUpload && dataType == DVR_FLOAT64) {
    setError("dvr_field_create: deferred upload does not support FLOAT64");
    return DVR_ERR_UNSUPPORTED;
  }
  const size_t esz = elementSize(dataType);
  if (esz == 0) {
    setError("dvr_field_create: invalid data type (StructuredRegularField.cpp:45-60)");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (gdims[0] == 0 || gdims[1] == 0 || gdims[2] == 0 || zEnd > gdims[2] || zBegin >= zEnd) {
    setError("dvr_field_create: bad dimensions / slab range");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (dvr_device_count() <= 0) {
    setError("dvr_field_create: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const bool whole = zBegin == 0 && zEnd == gdims[2];
  const uint32_t zTexBegin = zBegin > 0 ? zBegin - 1 : 0;
  const uint32_t zTexEnd = zEnd < gdims[2] ? zEnd + 1 : gdims[2]; // exclusive; one ghost slice each side
  const uint32_t texDepth = zTexEnd - zTexBegin;
  const size_t nvox = (size_t)gdims[0] * gdims[1] * texDepth;

  auto *f = new DvrField();
  cudaGetDevice(&f->device);

  // FLOAT64 is not a texturable format: convert to f32 first (the reference would fail here)
  const void *src = data;
  float *converted = nullptr;
  int texType = dataType;
  size_t texEsz = esz;
  if (dataType == DVR_FLOAT64) {
    double *staged = nullptr;
    const double *dsrc = (const double *)data;
    if (!dataIsDevice) {
      if (cudaMalloc(&staged, nvox * 8) != cudaSuccess) {
        delete f;
        return cudaFail(cudaGetLastError(), "cudaMalloc(f64 staging)");
      }
      cudaMemcpyAsync(staged, data, nvox * 8, cudaMemcpyHostToDevice, s);
      dsrc = staged;
    }
    if (cudaMalloc(&converted, nvox * 4) != cudaSuccess) {
      cudaFree(staged);
      delete f;
      return cudaFail(cudaGetLastError(), "cudaMalloc(f64->f32)");
    }
    launchConvertToFloat(dsrc, DVR_FLOAT64, converted, nvox, s);
    cudaStreamSynchronize(s);
    cudaFree(staged);
    src = converted;
    dataIsDevice = 1;
    texType = DVR_FLOAT32;
    texEsz = 4;
  }

  cudaChannelFormatKind kind = cudaChannelFormatKindFloat;
  if (texType == DVR_UFIXED8 || texType == DVR_UFIXED16)
    kind = cudaChannelFormatKindUnsigned;
  else if (texType == DVR_FIXED8 || texType == DVR_FIXED16)
    kind = cudaChannelFormatKindSigned;
  const cudaChannelFormatDesc desc = cudaCreateChannelDesc((int)texEsz * 8, 0, 0, 0, kind);
  // whole f32 fields get a surface view as well: device-memory input is then uploaded by the macrocell pass itself
  static const bool surfaceUpload = [] {
    const char *v = std::getenv("DVR_B200_SURFACE_UPLOAD");
    return !(v && v[0] == '0');
  }();
  const bool wantSurface = surfaceUpload && whole && dataType == DVR_FLOAT32;
  cudaError_t e = cudaMalloc3DArray(&f->array, &desc, make_cudaExtent(gdims[0], gdims[1], texDepth),
      wantSurface ? cudaArraySurfaceLoadStore : cudaArrayDefault);
  if (e != cudaSuccess) {
    cudaFree(converted);
    delete f;
    return cudaFail(e, "cudaMalloc3DArray");
  }
  f->texElementSize = texEsz;
  if (wantSurface) {
    cudaResourceDesc sd;
    std::memset(&sd, 0, sizeof(sd));
    sd.resType = cudaResourceTypeArray;
    sd.res.array.array = f->array;
    if (cudaCreateSurfaceObject(&f->surf, &sd) != cudaSuccess) {
      cudaGetLastError();
      f->surf = 0;
    }
  }
  const bool fusedUpload = !deferredUpload && f->surf && dataIsDevice
      && macrocellLinearIsVectorisable(data, make_int3((int)gdims[0], (int)gdims[1], (int)gdims[2]));
  if (!deferredUpload && !fusedUpload) {
    cudaMemcpy3DParms cp;
    std::memset(&cp, 0, sizeof(cp));
    cp.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), gdims[0] * texEsz, gdims[0], gdims[1]);
    cp.dstArray = f->array;
    cp.extent = make_cudaExtent(gdims[0], gdims[1], texDepth);
    cp.kind = dataIsDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    e = cudaMemcpy3DAsync(&cp, s);
    if (e == cudaSuccess && (!dataIsDevice || converted))
      e = cudaStreamSynchronize(s); // the caller's host buffer / our temporary may go away
  }
  cudaFree(converted);
  if (e != cudaSuccess) {
    if (f->surf)
      cudaDestroySurfaceObject(f->surf);
    cudaFreeArray(f->array);
    delete f;
    return cudaFail(e, "cudaMemcpy3D");
  }
  f->voxelBytes = nvox * texEsz;

  cudaResourceDesc rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = f->array;
  const bool isFloat = kind == cudaChannelFormatKindFloat;
  cudaTextureDesc td;
  std::memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = filter == DVR_FILTER_NEAREST ? cudaFilterModePoint : cudaFilterModeLinear;
  td.readMode = isFloat ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
  td.normalizedCoords = 1;
  e = cudaCreateTextureObject(&f->tex, &rd, &td, nullptr);
  if (e == cudaSuccess) {
    td.filterMode = cudaFilterModePoint;
    td.normalizedCoords = 0;
    e = cudaCreateTextureObject(&f->pointTex, &rd, &td, nullptr);
  }
  if (e != cudaSuccess) {
    dvr_field_destroy(f);
    return cudaFail(e, "cudaCreateTextureObject");
  }

  FieldDev &d = f->dev;
  d.tex = f->tex;
  f->dataType = dataType;
  setFieldGeometry(d, gdims, origin, spacing);
  d.zOwnBegin = whole ? 0 : (int)zBegin;
  d.zOwnEnd = whole ? (int)gdims[2] : (int)zEnd;
  d.zTexBegin = (int)zTexBegin;
  d.texDepth = (int)texDepth;
  f->ownLimitBegin = d.zOwnBegin;
  f->ownLimitEnd = d.zOwnEnd;
  d.gridDims = make_int3((int)((gdims[0] + 15) / 16), (int)((gdims[1] + 15) / 16), (int)((gdims[2] + 15) / 16));
  f->nCells = (size_t)d.gridDims.x * d.gridDims.y * d.gridDims.z;
  e = cudaMalloc(&f->ranges, f->nCells * sizeof(float2));
  if (e != cudaSuccess) {
    dvr_field_destroy(f);
    return cudaFail(e, "cudaMalloc(macrocell ranges)");
  }
  d.valueRanges = f->ranges;
  int rc = DVR_OK;
  if (!deferredUpload) {
    // whole f32 field handed over as linear device memory (ANARI_NV_ARRAY_CUDA, in-situ updates): separable build
    // straight from that memory; everything else reads the array through its point-sampled view
    if (dataIsDevice && dataType == DVR_FLOAT32 && whole)
      rc = launchMacrocellBuildLinear((const float *)data, d.dims, d.gridDims, f->ranges, fusedUpload ? f->surf : 0,
          (cudaStream_t)stream);
    else
      rc = dvr_field_build_macrocells(f, stream);
  }
  if (rc != DVR_OK) {
    dvr_field_destroy(f);
    return rc;
  }
  *out = f;
  return DVR_OK;
}

int dvr_field_create_structured(const void *data, int dataIsDevice, int dataType, const uint32_t dims[3],
    const float origin[3], const float spacing[3], int filter, void *stream, DvrField **out)
{
  if (!dims || !data) {
    setError("dvr_field_create_structured: null data / dims");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return createFieldImpl(data, dataIsDevice, dataType, dims, 0, dims[2], origin, spacing, filter, stream, out);
}

int dvr_field_create_structured_slab(const void *data, int dataIsDevice, int dataType, const uint32_t globalDims[3],
    uint32_t zBegin, uint32_t zEnd, const float origin[3], const float spacing[3], int filter, void *stream,
    DvrField **out)
{
  if (!globalDims) {
    setError("dvr_field_create_structured_slab: null dims");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return createFieldImpl(data, dataIsDevice, dataType, globalDims, zBegin, zEnd, origin, spacing, filter, stream, out);
}

// ---- NanoVDB ---------------------------------------------------------------------------------------------
extern "C++" {
namespace {
template <typename T>
T rd(const uint8_t *p, size_t off)
{
  T v;
  std::memcpy(&v, p + off, sizeof(T));
  return v;
}
} // namespace
}

// Gathers the grid into apron bricks (NvdbDev::brickTable) when table and bricks fit the budget: table <= 256 MiB,
// bricks <= min(4 GiB, a quarter of the free device memory).  Any failure leaves the field on the tree walk.
static void buildNvdbBricks(DvrField *f, const int bmin[3], const int bmax[3], cudaStream_t s)
{
  const char *env = std::getenv("DVR_B200_NVDB_BRICKS");
  if (env && env[0] == '0')
    return;
  if (bmax[0] < bmin[0] || bmax[1] < bmin[1] || bmax[2] < bmin[2])
    return; // empty grid
  FieldDev &d = f->dev;
  // cells of every base voxel whose stencil can touch the bounding box: [bmin - 1, bmax]
  const int3 org = make_int3((bmin[0] - 1) >> 3, (bmin[1] - 1) >> 3, (bmin[2] - 1) >> 3);
  const long long dx = (long long)(bmax[0] >> 3) - org.x + 1, dy = (long long)(bmax[1] >> 3) - org.y + 1,
                  dz = (long long)(bmax[2] >> 3) - org.z + 1;
  const long long cells = dx * dy * dz;
  if (dx <= 0 || dy <= 0 || dz <= 0 || cells > (256ll << 20) / (long long)sizeof(int2))
    return;
  const int3 dims = make_int3((int)dx, (int)dy, (int)dz);
  unsigned int *counter = nullptr;
  int2 *table = nullptr;
  float *bricks = nullptr;
  bool ok = cudaMalloc(&table, (size_t)cells * sizeof(int2)) == cudaSuccess
      && cudaMalloc(&counter, sizeof(unsigned int)) == cudaSuccess
      && cudaMemsetAsync(counter, 0, sizeof(unsigned int), s) == cudaSuccess
      && launchNvdbBrickBuild(d.nv, d.kind == FIELD_NANOVDB_QUANT, org, dims, table, nullptr, counter, 0u, s) == DVR_OK;
  unsigned int need = 0;
  ok = ok && cudaMemcpyAsync(&need, counter, sizeof(need), cudaMemcpyDeviceToHost, s) == cudaSuccess
      && cudaStreamSynchronize(s) == cudaSuccess;
  if (ok) {
    size_t freeB = 0, totalB = 0;
    cudaMemGetInfo(&freeB, &totalB);
    const size_t bytes = (size_t)std::max(need, 1u) * kNvdbBrickVoxels * sizeof(float);
    ok = bytes <= std::min<size_t>((size_t)4 << 30, freeB / 4) && cudaMalloc(&bricks, bytes) == cudaSuccess
        && cudaMemsetAsync(counter, 0, sizeof(unsigned int), s) == cudaSuccess
        && launchNvdbBrickBuild(d.nv, d.kind == FIELD_NANOVDB_QUANT, org, dims, table, bricks, counter, need, s) == DVR_OK
        && cudaStreamSynchronize(s) == cudaSuccess;
    if (ok) {
      f->brickTable = table;
      f->bricks = bricks;
      f->brickBytes = (size_t)cells * sizeof(int2) + bytes;
      d.nv.brickTable = table;
      d.nv.bricks = bricks;
      d.nv.brickOrg = org;
      d.nv.brickDims = dims;
      table = nullptr;
      bricks = nullptr;
    }
  }
  cudaGetLastError(); // a failed allocation is not an error of the field
  cudaFree(counter);
  cudaFree(table);
  cudaFree(bricks);
}

int dvr_field_create_nanovdb(const void *gridData, size_t bytes, int dataIsDevice, void *stream, DvrField **out)
{
  if (!gridData || !out || bytes < 672 + 64 + 64) {
    setError("dvr_field_create_nanovdb: null argument or buffer smaller than a NanoVDB grid header");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (dvr_device_count() <= 0) {
    setError("dvr_field_create_nanovdb: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  // header to the host: GridData (672 B) + TreeData (64 B)
  uint8_t head[736];
  if (dataIsDevice)
    DVR_CUDA(cudaMemcpy(head, gridData, sizeof(head), cudaMemcpyDeviceToHost));
  else
    std::memcpy(head, gridData, sizeof(head));
  const uint64_t magic = rd<uint64_t>(head, 0);
  const uint64_t kMagicNumb = 0x304244566f6e614eull, kMagicGrid = 0x314244566f6e614eull;
  if (magic != kMagicNumb && magic != kMagicGrid) {
    setError("dvr_field_create_nanovdb: not a NanoVDB grid buffer (bad magic)");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (rd<uint32_t>(head, 28) != 1u) {
    setError("dvr_field_create_nanovdb: the buffer must hold a single grid (NvdbRegularField.cpp:99-103)");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  const uint64_t gridSize = rd<uint64_t>(head, 32);
  if (gridSize > bytes) {
    setError("dvr_field_create_nanovdb: grid size exceeds the buffer");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  const uint32_t gridType = rd<uint32_t>(head, 636);
  // nanovdb::GridType: Float = 1, Fp4 = 13, Fp8 = 14, Fp16 = 15, FpN = 16 — the five types the reference's
  // marcher dispatches on (gpu/volumeIntegration.h:128-159); anything else renders nothing there.
  int codecLog2Bits = 0;
  bool quant = true;
  switch (gridType) {
  case 1u: quant = false; break;
  case 13u: codecLog2Bits = 2; break;
  case 14u: codecLog2Bits = 3; break;
  case 15u: codecLog2Bits = 4; break;
  case 16u: codecLog2Bits = -1; break;
  default:
    setError("dvr_field_create_nanovdb: unsupported GridType (Float, Fp4, Fp8, Fp16 and FpN are)");
    return DVR_ERR_UNSUPPORTED;
  }
  const int64_t rootOff = 672 + rd<int64_t>(head, 672 + 24); // TreeData::mNodeOffset[3]
  if (rootOff < 736 || (uint64_t)rootOff + 64 > gridSize) {
    setError("dvr_field_create_nanovdb: corrupt tree offsets");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  uint8_t root[64];
  if (dataIsDevice)
    DVR_CUDA(cudaMemcpy(root, (const uint8_t *)gridData + rootOff, sizeof(root), cudaMemcpyDeviceToHost));
  else
    std::memcpy(root, (const uint8_t *)gridData + rootOff, sizeof(root));

  // The device tree walk and the brick gather follow offsets stored in the buffer: validate all of them once on the
  // host (a device-resident grid is staged for the check) before anything is uploaded.
  {
    std::vector<uint8_t> staged;
    const uint8_t *hostGrid = (const uint8_t *)gridData;
    if (dataIsDevice) {
      staged.resize(gridSize);
      DVR_CUDA(cudaMemcpy(staged.data(), gridData, gridSize, cudaMemcpyDeviceToHost));
      hostGrid = staged.data();
    }
    const int bad = dvr_nvdb_validate_tree(hostGrid, gridSize);
    if (bad != 0) {
      setError("dvr_field_create_nanovdb: corrupt NanoVDB tree (node offsets leave the grid buffer, code "
          + std::to_string(bad) + ")");
      return DVR_ERR_INVALID_ARGUMENT;
    }
  }

  auto *f = new DvrField();
  cudaGetDevice(&f->device);
  cudaError_t e = cudaMalloc(&f->nvdbBlob, gridSize);
  if (e != cudaSuccess) {
    delete f;
    return cudaFail(e, "cudaMalloc(nanovdb grid)");
  }
  e = cudaMemcpyAsync(f->nvdbBlob, gridData, gridSize, dataIsDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && !dataIsDevice)
    e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    dvr_field_destroy(f);
    return cudaFail(e, "cudaMemcpy(nanovdb grid)");
  }
  f->voxelBytes = gridSize;

  FieldDev &d = f->dev;
  std::memset(&d, 0, sizeof(d));
  d.kind = quant ? FIELD_NANOVDB_QUANT : FIELD_NANOVDB;
  d.nv.codecLog2Bits = codecLog2Bits;
  d.nv.root = (const uint8_t *)f->nvdbBlob + rootOff;
  d.nv.tileCount = rd<uint32_t>(root, 24);
  d.nv.background = rd<float>(root, 28);
  for (int i = 0; i < 9; ++i)
    d.nv.invMat[i] = rd<float>(head, 296 + 36 + 4 * i);
  for (int i = 0; i < 3; ++i)
    d.nv.vec[i] = rd<float>(head, 296 + 72 + 4 * i);
  const int bmin[3] = {rd<int32_t>(root, 0), rd<int32_t>(root, 4), rd<int32_t>(root, 8)};
  const int bmax[3] = {rd<int32_t>(root, 12), rd<int32_t>(root, 16), rd<int32_t>(root, 20)};
  d.nv.bboxMin = make_int3(bmin[0], bmin[1], bmin[2]);
  d.dims = make_int3(std::max(bmax[0] - bmin[0] + 1, 1), std::max(bmax[1] - bmin[1] + 1, 1),
      std::max(bmax[2] - bmin[2] + 1, 1));
  // bounds = worldBBox, voxelSize -> step (NvdbRegularField.cpp:105-113,124-127)
  double wb[6], vs[3];
  for (int i = 0; i < 6; ++i)
    wb[i] = rd<double>(head, 560 + 8 * i);
  for (int i = 0; i < 3; ++i)
    vs[i] = rd<double>(head, 608 + 8 * i);
  d.boundsLo = make_float3((float)wb[0], (float)wb[1], (float)wb[2]);
  d.boundsHi = make_float3((float)wb[3], (float)wb[4], (float)wb[5]);
  d.origin = d.boundsLo;
  d.spacing = make_float3((float)vs[0], (float)vs[1], (float)vs[2]);
  d.invSpacing = make_float3(1.f / d.spacing.x, 1.f / d.spacing.y, 1.f / d.spacing.z);
  d.stepSize = std::fmin(std::fmin(d.spacing.x, d.spacing.y), d.spacing.z) / 2.0f;
  d.zOwnBegin = 0;
  d.zOwnEnd = d.dims.z;
  d.zTexBegin = 0;
  d.texDepth = d.dims.z;
  d.gridDims = make_int3((d.dims.x + 15) / 16, (d.dims.y + 15) / 16, (d.dims.z + 15) / 16);
  f->nCells = (size_t)d.gridDims.x * d.gridDims.y * d.gridDims.z;
  e = cudaMalloc(&f->ranges, f->nCells * sizeof(float2));
  if (e != cudaSuccess) {
    dvr_field_destroy(f);
    return cudaFail(e, "cudaMalloc(macrocell ranges)");
  }
  d.valueRanges = f->ranges;
  const int rc = dvr_field_build_macrocells(f, stream);
  if (rc != DVR_OK) {
    dvr_field_destroy(f);
    return rc;
  }
  buildNvdbBricks(f, bmin, bmax, s);
  *out = f;
  return DVR_OK;
}

int dvr_field_set_owned_slices(DvrField *f, uint32_t zBegin, uint32_t zEnd)
{
  if (!f) {
    setError("dvr_field_set_owned_slices: null field");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (f->dev.kind != FIELD_STRUCTURED) {
    setError("dvr_field_set_owned_slices: structuredRegular slab fields only");
    return DVR_ERR_UNSUPPORTED;
  }
  if (zBegin >= zEnd || (int)zBegin < f->ownLimitBegin || (int)zEnd > f->ownLimitEnd) {
    setError("dvr_field_set_owned_slices: [" + std::to_string(zBegin) + ", " + std::to_string(zEnd)
        + ") is empty or leaves the range the slab was created with [" + std::to_string(f->ownLimitBegin) + ", "
        + std::to_string(f->ownLimitEnd) + ")");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  f->dev.zOwnBegin = (int)zBegin;
  f->dev.zOwnEnd = (int)zEnd;
  return DVR_OK;
}

int dvr_field_owned_slices(const DvrField *f, uint32_t *zBegin, uint32_t *zEnd, uint32_t *limitBegin, uint32_t *limitEnd)
{
  if (!f) {
    setError("dvr_field_owned_slices: null field");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (zBegin) *zBegin = (uint32_t)f->dev.zOwnBegin;
  if (zEnd) *zEnd = (uint32_t)f->dev.zOwnEnd;
  if (limitBegin) *limitBegin = (uint32_t)f->ownLimitBegin;
  if (limitEnd) *limitEnd = (uint32_t)f->ownLimitEnd;
  return DVR_OK;
}

int dvr_field_upload_slices(DvrField *f, const void *data, int dataIsDevice, uint32_t firstResidentSlice,
    uint32_t nSlices, void *stream)
{
  if (!f || f->dev.kind != FIELD_STRUCTURED || !data || nSlices == 0
      || firstResidentSlice + nSlices > (uint32_t)f->dev.texDepth) {
    setError("dvr_field_upload_slices: bad argument / slice range");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemcpy3DParms cp;
  std::memset(&cp, 0, sizeof(cp));
  cp.srcPtr = make_cudaPitchedPtr(const_cast<void *>(data), f->dev.dims.x * f->texElementSize, f->dev.dims.x, f->dev.dims.y);
  cp.dstArray = f->array;
  cp.dstPos = make_cudaPos(0, 0, firstResidentSlice);
  cp.extent = make_cudaExtent(f->dev.dims.x, f->dev.dims.y, nSlices);
  cp.kind = dataIsDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  DVR_CUDA(cudaMemcpy3DAsync(&cp, s));
  if (!dataIsDevice)
    DVR_CUDA(cudaStreamSynchronize(s));
  return DVR_OK;
}

int dvr_field_update_structured(DvrField *f, const void *data, int dataIsDevice, int dataType, const float origin[3],
    const float spacing[3], void *stream)
{
  if (!f || !data || !origin || !spacing) {
    setError("dvr_field_update_structured: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  const FieldDev &d0 = f->dev;
  if (d0.kind != FIELD_STRUCTURED || d0.zTexBegin != 0 || d0.texDepth != d0.dims.z || dataType != f->dataType
      || dataType == DVR_FLOAT64) {
    setError("dvr_field_update_structured: only a whole structuredRegular field of unchanged element type (not "
             "FLOAT64) updates in place; destroy and create instead");
    return DVR_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const uint32_t gdims[3] = {(uint32_t)d0.dims.x, (uint32_t)d0.dims.y, (uint32_t)d0.dims.z};
  setFieldGeometry(f->dev, gdims, origin, spacing);
  const bool linear = dataIsDevice && dataType == DVR_FLOAT32;
  const bool fused = linear && f->surf && macrocellLinearIsVectorisable(data, f->dev.dims);
  if (!fused) {
    const int up = dvr_field_upload_slices(f, data, dataIsDevice, 0, gdims[2], stream);
    if (up != DVR_OK)
      return up;
  }
  if (linear)
    return launchMacrocellBuildLinear((const float *)data, f->dev.dims, f->dev.gridDims, f->ranges,
        fused ? f->surf : 0, s);
  return dvr_field_build_macrocells(f, stream);
}

int dvr_field_destroy(DvrField *f)
{
  if (!f)
    return DVR_OK;
  if (f->tex) cudaDestroyTextureObject(f->tex);
  if (f->pointTex) cudaDestroyTextureObject(f->pointTex);
  if (f->surf) cudaDestroySurfaceObject(f->surf);
  if (f->array) cudaFreeArray(f->array);
  if (f->ranges) cudaFree(f->ranges);
  if (f->nvdbBlob) cudaFree(f->nvdbBlob);
  if (f->brickTable) cudaFree(f->brickTable);
  if (f->bricks) cudaFree(f->bricks);
  delete f;
  return DVR_OK;
}

int dvr_field_bounds(const DvrField *f, float lower[3], float upper[3])
{
  if (!f || !lower || !upper) {
    setError("dvr_field_bounds: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  st3(lower, f->dev.boundsLo);
  st3(upper, f->dev.boundsHi);
  return DVR_OK;
}

int dvr_field_step_size(const DvrField *f, float *stepSize)
{
  if (!f || !stepSize) {
    setError("dvr_field_step_size: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  *stepSize = f->dev.stepSize;
  return DVR_OK;
}

int dvr_field_device_bytes(const DvrField *f, size_t *bytes)
{
  if (!f || !bytes) {
    setError("dvr_field_device_bytes: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  *bytes = f->voxelBytes + f->nCells * sizeof(float2) + f->brickBytes;
  return DVR_OK;
}

int dvr_field_build_macrocells(DvrField *f, void *stream)
{
  if (!f) {
    setError("dvr_field_build_macrocells: null field");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (f->dev.kind >= FIELD_NANOVDB)
    return launchMacrocellBuildNvdb(f->dev, f->ranges, (cudaStream_t)stream);
  return launchMacrocellBuild(
      f->pointTex, f->dev.dims, f->dev.zTexBegin, f->dev.texDepth, f->dev.gridDims, f->ranges, (cudaStream_t)stream);
}

int dvr_field_macrocells(const DvrField *f, uint32_t gridDims[3], const float **valueRangesDev)
{
  if (!f) {
    setError("dvr_field_macrocells: null field");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (gridDims) {
    gridDims[0] = f->dev.gridDims.x;
    gridDims[1] = f->dev.gridDims.y;
    gridDims[2] = f->dev.gridDims.z;
  }
  if (valueRangesDev)
    *valueRangesDev = (const float *)f->ranges;
  return DVR_OK;
}

int dvr_field_value_range(const DvrField *f, void *stream, float range[2])
{
  if (!f || !range) {
    setError("dvr_field_value_range: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  float2 *d = nullptr;
  DVR_CUDA(cudaMalloc(&d, sizeof(float2)));
  int rc = launchRangeReduce(f->ranges, f->nCells, d, (cudaStream_t)stream);
  float2 h = make_float2(0.f, 1.f);
  if (rc == DVR_OK) {
    cudaError_t e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess)
      rc = cudaFail(e, "dvr_field_value_range copy");
  }
  cudaFree(d);
  range[0] = h.x;
  range[1] = h.y;
  return rc;
}

// ---- background image --------------------------------------------------------------------------------

// convertComponentUint8, utility/CudaImageTexture.cpp:43-58 (the float -> uint8 conversions truncate; values outside
// [0, 255] are undefined behaviour there and saturate here)
static inline uint8_t toU8(float v)
{
  if (!(v > 0.f))
    return 0;
  return v >= 255.f ? (uint8_t)255 : (uint8_t)v;
}

int dvr_image_create(const void *pixels, int componentType, int channels, uint32_t width, uint32_t height,
    void *stream, DvrImage **out)
{
  if (!pixels || !out || width == 0 || height == 0 || channels < 1 || channels > 4
      || componentType < DVR_IMAGE_FLOAT32 || componentType > DVR_IMAGE_SRGB8) {
    setError("dvr_image_create: bad argument (host pixels, 1..4 channels, non-empty size)");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (dvr_device_count() <= 0) {
    setError("dvr_image_create: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  const size_t n = (size_t)width * height;
  const int nc = channels == 3 ? 4 : channels; // three channels are padded with alpha 255
  std::vector<uint8_t> staging(n * (size_t)nc);
  size_t o = 0;
  for (size_t i = 0; i < n * (size_t)channels; ++i) {
    uint8_t c;
    switch (componentType) {
    case DVR_IMAGE_FLOAT32: c = toU8(((const float *)pixels)[i] * 255); break;
    case DVR_IMAGE_UFIXED16: c = toU8((((const uint16_t *)pixels)[i] / 65535.f) * 255); break;
    case DVR_IMAGE_UFIXED32: c = toU8((((const uint32_t *)pixels)[i] / float(UINT32_MAX)) * 255); break;
    case DVR_IMAGE_SRGB8: {
      // glm::convertSRGBToLinear (gtc/color_space.inl:33-43) on every component, alpha included, like the reference
      const float v = ((const uint8_t *)pixels)[i] / 255.f;
      const float lin = v <= 0.04045f ? v * 0.07739938080495356037151702786378f
                                      : std::pow((v + 0.055f) * 0.94786729857819905213270142180095f, 2.4f);
      c = toU8(lin * 255);
      break;
    }
    default: c = ((const uint8_t *)pixels)[i]; break;
    }
    staging[o++] = c;
    if (channels == 3 && o % 4 == 3)
      staging[o++] = 255;
  }
  auto *img = new DvrImage();
  img->width = width;
  img->height = height;
  img->channels = nc;
  const cudaChannelFormatDesc desc =
      cudaCreateChannelDesc(8, nc >= 2 ? 8 : 0, nc >= 3 ? 8 : 0, nc >= 4 ? 8 : 0, cudaChannelFormatKindUnsigned);
  cudaError_t e = cudaMallocArray(&img->array, &desc, width, height);
  cudaStream_t s = (cudaStream_t)stream;
  if (e == cudaSuccess)
    e = cudaMemcpy2DToArrayAsync(img->array, 0, 0, staging.data(), (size_t)width * nc, (size_t)width * nc, height,
        cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(s); // the staging buffer is pageable and local
  if (e == cudaSuccess) {
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = img->array;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    e = cudaCreateTextureObject(&img->tex, &rd, &td, nullptr);
  }
  if (e != cudaSuccess) {
    dvr_image_destroy(img);
    return cudaFail(e, "dvr_image_create");
  }
  *out = img;
  return DVR_OK;
}

int dvr_image_destroy(DvrImage *img)
{
  if (!img)
    return DVR_OK;
  if (img->tex) cudaDestroyTextureObject(img->tex);
  if (img->array) cudaFreeArray(img->array);
  delete img;
  return DVR_OK;
}

// ---- volumes -----------------------------------------------------------------------------------------

int dvr_volume_update(DvrVolume *v, const float *tfRgba, const float valueRange[2], float unitDistance, uint32_t id,
    void *stream)
{
  if (!v || !tfRgba || !valueRange) {
    setError("dvr_volume_update: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  {
    const int3 g = v->field->dev.gridDims;
    if (v->field->nCells != v->allocCells || g.x != v->allocGridDims.x || g.y != v->allocGridDims.y
        || g.z != v->allocGridDims.z) {
      setError("dvr_volume_update: the field's macrocell grid differs from the one this volume was created for; "
               "destroy and create instead");
      return DVR_ERR_UNSUPPORTED;
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  DVR_CUDA(cudaMemcpyAsync(v->tf, tfRgba, DVR_TF_SIZE * sizeof(float4), cudaMemcpyHostToDevice, s));
  DVR_CUDA(cudaStreamSynchronize(s)); // tfRgba is caller-owned pageable memory
  v->vrLo = valueRange[0];
  v->vrHi = valueRange[1];
  v->oneOverUnitDistance = 1.0f / unitDistance;
  v->id = id;
  v->ddaValid = false;
  const int rc = launchMajorants(v->field->ranges, v->field->nCells, v->tf, v->vrLo, v->vrHi, v->maxOpacities, s);
  if (rc != DVR_OK)
    return rc;
  const int rc2 = launchMajorantsCoarse(v->maxOpacities, v->field->dev.gridDims, v->maxOpacitiesCoarse, v->coarseDims, s);
  if (rc2 != DVR_OK)
    return rc2;
  const int rc3 = launchMajorantsCoarse(v->maxOpacitiesCoarse, v->coarseDims, v->maxOpacitiesCoarse2, v->coarse2Dims, s);
  if (rc3 != DVR_OK)
    return rc3;
  // how much of the volume skipping could exploit (drives DVR_SKIP_AUTO)
  unsigned long long *dCount = nullptr, hCount = 0;
  DVR_CUDA(scratchAllocAsync((void **)&dCount, sizeof(unsigned long long), s));
  DVR_CUDA(cudaMemsetAsync(dCount, 0, sizeof(unsigned long long), s));
  const int rc4 = launchCountEmpty(v->maxOpacities, v->field->nCells, dCount, s);
  DVR_CUDA(cudaMemcpyAsync(&hCount, dCount, sizeof(hCount), cudaMemcpyDeviceToHost, s));
  DVR_CUDA(cudaFreeAsync(dCount, s));
  DVR_CUDA(cudaStreamSynchronize(s));
  v->emptyFraction = v->field->nCells ? (float)((double)hCount / (double)v->field->nCells) : 0.f;
  return rc4;
}

int dvr_volume_create(const DvrField *field, const float *tfRgba, const float valueRange[2], float unitDistance,
    uint32_t id, void *stream, DvrVolume **out)
{
  if (!field || !tfRgba || !valueRange || !out) {
    setError("dvr_volume_create: null argument (missing parameter 'value' on transferFunction1D?)");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  auto *v = new DvrVolume();
  v->field = field;
  cudaError_t e = cudaMalloc(&v->tf, DVR_TF_SIZE * sizeof(float4));
  if (e == cudaSuccess)
    e = cudaMalloc(&v->maxOpacities, field->nCells * sizeof(float));
  const int3 g = field->dev.gridDims;
  v->allocCells = field->nCells;
  v->allocGridDims = g;
  v->coarseDims = make_int3((g.x + 3) / 4, (g.y + 3) / 4, (g.z + 3) / 4);
  if (e == cudaSuccess)
    e = cudaMalloc(&v->maxOpacitiesCoarse, (size_t)v->coarseDims.x * v->coarseDims.y * v->coarseDims.z * sizeof(float));
  v->coarse2Dims = make_int3((v->coarseDims.x + 3) / 4, (v->coarseDims.y + 3) / 4, (v->coarseDims.z + 3) / 4);
  if (e == cudaSuccess)
    e = cudaMalloc(&v->maxOpacitiesCoarse2,
        (size_t)v->coarse2Dims.x * v->coarse2Dims.y * v->coarse2Dims.z * sizeof(float));
  if (e != cudaSuccess) {
    dvr_volume_destroy(v);
    return cudaFail(e, "cudaMalloc(volume)");
  }
  const int rc = dvr_volume_update(v, tfRgba, valueRange, unitDistance, id, stream);
  if (rc != DVR_OK) {
    dvr_volume_destroy(v);
    return rc;
  }
  *out = v;
  return DVR_OK;
}

int dvr_volume_destroy(DvrVolume *v)
{
  if (!v)
    return DVR_OK;
  if (v->tf) cudaFree(v->tf);
  if (v->maxOpacities) cudaFree(v->maxOpacities);
  if (v->maxOpacitiesCoarse) cudaFree(v->maxOpacitiesCoarse);
  if (v->maxOpacitiesCoarse2) cudaFree(v->maxOpacitiesCoarse2);
  if (v->ddaMaxOpacities) cudaFree(v->ddaMaxOpacities);
  if (v->ddaRanges) cudaFree(v->ddaRanges);
  delete v;
  return DVR_OK;
}

static int ensureDdaGrid(DvrVolume *v, bool referenceBuild, cudaStream_t s);

int dvr_volume_dda_majorants(DvrVolume *v, int32_t referenceBuild, void *stream, uint32_t dims[3],
    const float **maxOpacitiesDev)
{
  if (!v || !dims || !maxOpacitiesDev) {
    setError("dvr_volume_dda_majorants: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  const int rc = ensureDdaGrid(v, referenceBuild != 0, (cudaStream_t)stream);
  if (rc != DVR_OK)
    return rc;
  dims[0] = (uint32_t)v->field->dev.gridDims.x;
  dims[1] = (uint32_t)v->field->dev.gridDims.y;
  dims[2] = (uint32_t)v->field->dev.gridDims.z;
  *maxOpacitiesDev = v->ddaMaxOpacities;
  return DVR_OK;
}

int dvr_volume_majorants(const DvrVolume *v, const float **maxOpacitiesDev)
{
  if (!v || !maxOpacitiesDev) {
    setError("dvr_volume_majorants: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  *maxOpacitiesDev = v->maxOpacities;
  return DVR_OK;
}

// ---- render ------------------------------------------------------------------------------------------

static void fillCamera(const DvrCamera *c, CameraDev &d)
{
  d.type = c->type;
  d.region = make_float4(c->region[0], c->region[1], c->region[2], c->region[3]);
  d.pos = v3(c->pos);
  d.dir = v3(c->dir);
  d.du = v3(c->du);
  d.dv = v3(c->dv);
  d.p00 = v3(c->p00);
  d.scaledAperture = c->scaledAperture;
  d.aspect = c->aspect;
}

static void fillInstance(const DvrVolumeInstance &in, InstanceDev &d)
{
  const DvrVolume *v = in.volume;
  d.v.f = v->field->dev;
  d.v.tf = v->tf;
  d.v.vrLower = v->vrLo;
  d.v.vrUpper = v->vrHi;
  d.v.oneOverUnitDistance = v->oneOverUnitDistance;
  d.v.id = v->id;
  d.v.maxOpacities = v->maxOpacities;
  d.v.maxOpacitiesCoarse = v->maxOpacitiesCoarse;
  d.v.coarseDims = v->coarseDims;
  d.v.maxOpacitiesCoarse2 = v->maxOpacitiesCoarse2;
  d.v.coarse2Dims = v->coarse2Dims;
  d.v.ddaMaxOpacities = v->ddaMaxOpacities;
  d.v.ddaDims = v->field->dev.gridDims; // ceil(dims/16), UniformGrid.cu:152-154
  static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  std::memcpy(d.xfm, in.worldToObject, sizeof(d.xfm));
  d.instId = in.instanceId;
  d.identity = std::memcmp(in.worldToObject, ident, sizeof(ident)) == 0 ? 1u : 0u;
}

// UniformGrid::init/buildGrid/computeMaxOpacities for the delta tracker: the reference's grid geometry
// (gridDims cells dividing the bounds evenly) with conservative content.  Built lazily per volume.
static int ensureDdaGrid(DvrVolume *v, bool referenceBuild, cudaStream_t s)
{
  if (v->ddaValid && v->ddaReference == referenceBuild)
    return DVR_OK;
  const DvrField *f = v->field;
  const int3 g = f->dev.gridDims;
  const size_t n = f->nCells;
  if (!v->ddaRanges)
    DVR_CUDA(cudaMalloc(&v->ddaRanges, n * sizeof(float2)));
  if (!v->ddaMaxOpacities)
    DVR_CUDA(cudaMalloc(&v->ddaMaxOpacities, n * sizeof(float)));
  // voxel units spanned by the bounds: dims-1 for node-centred structured fields, dims for NanoVDB boxes
  const float3 span = f->dev.kind >= FIELD_NANOVDB
      ? make_float3((float)f->dev.dims.x, (float)f->dev.dims.y, (float)f->dev.dims.z)
      : make_float3((float)f->dev.dims.x - 1.f, (float)f->dev.dims.y - 1.f, (float)f->dev.dims.z - 1.f);
  const float3 w = make_float3(span.x / (float)g.x, span.y / (float)g.y, span.z / (float)g.z);
  int rc;
  if (referenceBuild)
    rc = launchReferenceGridBuild(f->dev, g, v->tf, v->ddaRanges, v->ddaMaxOpacities, s);
  else {
    rc = launchDdaRangeBuild(f->dev, f->pointTex, g, w, v->ddaRanges, s);
    if (rc != DVR_OK)
      return rc;
    rc = launchMajorants(v->ddaRanges, n, v->tf, v->vrLo, v->vrHi, v->ddaMaxOpacities, s);
  }
  if (rc == DVR_OK) {
    v->ddaValid = true;
    v->ddaReference = referenceBuild;
  }
  return rc;
}

// Conservative pixel rectangle that contains every pixel whose (jittered) primary ray can hit the instance's
// bounds: the projection of the 8 box corners through the camera model of cameraCreateRay, padded by 2 pixels.
// Returns false (=> whole frame) whenever the bound cannot be trusted: thin-lens cameras, a transformed instance,
// a corner behind the eye, a degenerate camera basis.
static bool screenRectOfBox(const DvrCamera *c, const float3 lo, const float3 hi, uint32_t W, uint32_t H, int rect[4]);

static bool screenRectOfBounds(const DvrCamera *c, const DvrVolumeInstance *in, uint32_t W, uint32_t H, int rect[4])
{
  static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  if (std::memcmp(in->worldToObject, ident, sizeof(ident)) != 0)
    return false;
  return screenRectOfBox(c, in->volume->field->dev.boundsLo, in->volume->field->dev.boundsHi, W, H, rect);
}

static bool screenRectOfBox(const DvrCamera *c, const float3 lo, const float3 hi, uint32_t W, uint32_t H, int rect[4])
{
  if (c->scaledAperture > 0.f)
    return false;
  const double rw = (double)c->region[2] - c->region[0], rh = (double)c->region[3] - c->region[1];
  if (!(std::fabs(rw) > 1e-12) || !(std::fabs(rh) > 1e-12))
    return false;
  const bool persp = c->type == DVR_CAMERA_PERSPECTIVE;
  // columns of the 3x3 system: du, dv, and p00 (perspective) or dir (orthographic)
  const double a[3] = {c->du[0], c->du[1], c->du[2]}, b[3] = {c->dv[0], c->dv[1], c->dv[2]};
  const double k[3] = {persp ? c->p00[0] : c->dir[0], persp ? c->p00[1] : c->dir[1], persp ? c->p00[2] : c->dir[2]};
  const double det = a[0] * (b[1] * k[2] - b[2] * k[1]) - b[0] * (a[1] * k[2] - a[2] * k[1]) + k[0] * (a[1] * b[2] - a[2] * b[1]);
  if (!(std::fabs(det) > 1e-30))
    return false;
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int i = 0; i < 8; ++i) {
    const double P[3] = {(i & 1) ? hi.x : lo.x, (i & 2) ? hi.y : lo.y, (i & 4) ? hi.z : lo.z};
    const double o[3] = {persp ? c->pos[0] : c->p00[0], persp ? c->pos[1] : c->p00[1], persp ? c->pos[2] : c->p00[2]};
    const double r[3] = {P[0] - o[0], P[1] - o[1], P[2] - o[2]};
    // Cramer: r = x*a + y*b + z*k
    const double x = (r[0] * (b[1] * k[2] - b[2] * k[1]) - b[0] * (r[1] * k[2] - r[2] * k[1]) + k[0] * (r[1] * b[2] - r[2] * b[1])) / det;
    const double y = (a[0] * (r[1] * k[2] - r[2] * k[1]) - r[0] * (a[1] * k[2] - a[2] * k[1]) + k[0] * (a[1] * r[2] - a[2] * r[1])) / det;
    const double z = (a[0] * (b[1] * r[2] - b[2] * r[1]) - b[0] * (a[1] * r[2] - a[2] * r[1]) + r[0] * (a[1] * b[2] - a[2] * b[1])) / det;
    double sx = x, sy = y;
    if (persp) {
      if (!(z > 1e-6)) // a corner beside or behind the eye: the projection is unbounded
        return false;
      sx = x / z;
      sy = y / z;
    }
    // undo the image-region mapping of cameraCreateRay (sx' = mix(region.x, region.z, sx))
    const double u = (sx - c->region[0]) / rw * (double)W, v = (sy - c->region[1]) / rh * (double)H;
    if (!(u == u) || !(v == v))
      return false;
    xmin = std::min(xmin, u);
    xmax = std::max(xmax, u);
    ymin = std::min(ymin, v);
    ymax = std::max(ymax, v);
  }
  const double pad = 2.0; // pixel jitter is < 1 pixel; the rest is rounding head-room
  rect[0] = (int)std::floor(std::max(xmin - pad, 0.0));
  rect[1] = (int)std::floor(std::max(ymin - pad, 0.0));
  rect[2] = (int)std::ceil(std::min(xmax + pad, (double)W));
  rect[3] = (int)std::ceil(std::min(ymax + pad, (double)H));
  return true;
}

static int renderImpl(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instances,
    uint32_t nInstances, const DvrFrameBuffers *b, DvrRenderStats *statsDev, bool stats, void *stream,
    const DvrSceneParams *scene = nullptr)
{
  if (!p || !camera || !b || (nInstances && !instances)) {
    setError("dvr_render: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (!b->colorAccumulation || !b->outColor) {
    setError("dvr_render: colorAccumulation and outColor buffers are required");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (p->width == 0 || p->height == 0 || p->numIterations < 1 || p->format < 0 || p->format > 2
      || p->checkerboardID > 3 || p->integrator < 0 || p->integrator > DVR_INTEGRATOR_TEST) {
    setError("dvr_render: invalid frame parameters");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  for (uint32_t i = 0; i < nInstances; ++i)
    if (!instances[i].volume || !instances[i].volume->field) {
      setError("dvr_render: instance without a valid volume");
      return DVR_ERR_INVALID_ARGUMENT;
    }
  if (dvr_device_count() <= 0) {
    setError("dvr_render: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  if (stats && !statsDev) {
    setError("dvr_render_instrumented: statsDev is null");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  // a world with surfaces takes the mixed-scene kernel (dvr_scene.cu); without any it is the plain volume frame
  const bool hasSurfaces = sceneHasSurfaces(scene);
  if (hasSurfaces && p->integrator != DVR_INTEGRATOR_RAYCAST && p->integrator != DVR_INTEGRATOR_DEFAULT) {
    setError("dvr_render_scene: surfaces are rendered by the raycast and default integrators only");
    return DVR_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;

  FrameLaunch L;
  std::memset(&L, 0, sizeof(L));
  L.width = p->width;
  L.height = p->height;
  L.invW = 1.f / (float)p->width; // Frame.cu:112
  L.invH = 1.f / (float)p->height;
  L.format = p->format;
  L.integrator = p->integrator;
  L.frameID = p->frameID;
  L.checkerboardID = p->checkerboardID;
  L.numIterations = p->checkerboardID >= 0 ? 1 : p->numIterations; // Renderer.cpp:168-169
  L.invSamplingRate = p->inverseVolumeSamplingRate;
  L.background = make_float4(p->background[0], p->background[1], p->background[2], p->background[3]);
  L.bgTex = p->backgroundImage ? p->backgroundImage->tex : 0;
  L.maxDepth = p->maxDepth <= 0 ? 5 : (p->maxDepth > 256 ? 256 : p->maxDepth);
  L.ambientIntensity = p->ambientRadiance;
  L.occlusionDistance = p->occlusionDistance > 0.f ? p->occlusionDistance : 1e20f;
  if (p->integrator == DVR_INTEGRATOR_DPT)
    for (uint32_t i = 0; i < nInstances; ++i) {
      const int rcg = ensureDdaGrid(const_cast<DvrVolume *>(instances[i].volume), p->dptReferenceGrid != 0, s);
      if (rcg != DVR_OK)
        return rcg;
    }
  L.tileRank = p->tileRanks > 1 ? p->tileRank : 0;
  L.tileRanks = p->tileRanks > 1 ? p->tileRanks : 1;
  L.tileBand = p->tileBand > 1 ? (uint32_t)p->tileBand : 1u;
  L.launchW = p->checkerboardID >= 0 ? (p->width + 1) / 2 : p->width; // Frame.cu:287-288
  L.launchH = p->checkerboardID >= 0 ? (p->height + 1) / 2 : p->height;
  L.tilesX = (L.launchW + kTileW - 1) / kTileW;
  L.tilesY = (L.launchH + kTileH - 1) / kTileH;
  fillCamera(camera, L.cam);
  L.fb.accum = (float4 *)b->colorAccumulation;
  L.fb.outU32 = (uint32_t *)b->outColor;
  L.fb.outF32 = (float4 *)b->outColor;
  L.fb.outMirror = b->outColorMirror;
  L.fb.depth = b->depth;
  L.fb.depthMirror = b->depth ? b->depthMirror : nullptr;
  L.fb.primId = b->primId;
  L.fb.objId = b->objId;
  L.fb.instId = b->instId;
  L.fb.albedo = b->albedo;
  L.fb.normal = b->normal;
  L.nInst = (int)nInstances;

  // background-only pixels without ray set-up (marching integrators; not when the ray direction is an output)
  L.missValid = 0;
  if ((p->integrator == DVR_INTEGRATOR_RAYCAST || p->integrator == DVR_INTEGRATOR_DEFAULT) && !b->normal
      && !hasSurfaces) {
    int u[4] = {(int)p->width, (int)p->height, 0, 0};
    bool ok = true;
    for (uint32_t i = 0; i < nInstances && ok; ++i) {
      int r[4];
      ok = screenRectOfBounds(camera, &instances[i], p->width, p->height, r);
      if (ok && r[2] > r[0] && r[3] > r[1]) {
        u[0] = std::min(u[0], r[0]);
        u[1] = std::min(u[1], r[1]);
        u[2] = std::max(u[2], r[2]);
        u[3] = std::max(u[3], r[3]);
      }
    }
    if (ok) {
      L.missValid = 1;
      L.missX0 = u[0];
      L.missY0 = u[1];
      L.missX1 = u[2];
      L.missY1 = u[3];
    }
  }
  // Tile window + background sweep: when the rectangle is known (and the launch grid is the pixel grid), the 8x4
  // tiles cover only the tile-aligned rectangle; everything outside is swept by a second, thin kernel in row-major
  // order (coalesced 512 B / 128 B stores for the accumulation and the mirrored colour buffer instead of 32-byte
  // tile rows) on an auxiliary stream, concurrently with the march.
  L.tileX0 = L.tileY0 = 0;
  L.tilesW = L.tilesX;
  L.tilesH = L.tilesY;
  L.sweepOutside = 0;
  if (L.missValid && p->checkerboardID < 0 && p->tileRanks <= 1 && !stats) {
    const int W = (int)p->width, H = (int)p->height;
    int x0 = std::min(std::max(L.missX0, 0), W), y0 = std::min(std::max(L.missY0, 0), H);
    int x1 = std::min(std::max(L.missX1, x0), W), y1 = std::min(std::max(L.missY1, y0), H);
    x0 = x0 / kTileW * kTileW;
    y0 = y0 / kTileH * kTileH;
    x1 = std::min((x1 + kTileW - 1) / kTileW * kTileW, (int)L.tilesX * kTileW);
    y1 = std::min((y1 + kTileH - 1) / kTileH * kTileH, (int)L.tilesY * kTileH);
    // worth a second launch only when a good part of the frame lies outside
    if ((size_t)(x1 - x0) * (size_t)(y1 - y0) * 4 <= (size_t)W * (size_t)H * 3) {
      L.missX0 = x0;
      L.missY0 = y0;
      L.missX1 = x1;
      L.missY1 = y1;
      L.tileX0 = (uint32_t)(x0 / kTileW);
      L.tileY0 = (uint32_t)(y0 / kTileH);
      L.tilesW = (uint32_t)((x1 - x0) / kTileW);
      L.tilesH = (uint32_t)((y1 - y0) / kTileH);
      L.sweepOutside = 1;
    }
  }
  bool skip = p->useMacrocellSkipping == DVR_SKIP_ON;
  if (p->useMacrocellSkipping == DVR_SKIP_AUTO) // images are identical either way: pick the faster kernel
    for (uint32_t i = 0; i < nInstances; ++i)
      skip = skip || instances[i].volume->emptyFraction >= kSkipAutoThreshold;
  if (nInstances <= (uint32_t)kMaxInlineInstances) {
    for (uint32_t i = 0; i < nInstances; ++i)
      fillInstance(instances[i], L.inl[i]);
  } else {
    std::vector<InstanceDev> tmp(nInstances);
    for (uint32_t i = 0; i < nInstances; ++i)
      fillInstance(instances[i], tmp[i]);
    InstanceDev *ext = nullptr;
    DVR_CUDA(scratchAllocAsync((void **)&ext, nInstances * sizeof(InstanceDev), s));
    DVR_CUDA(cudaMemcpyAsync(ext, tmp.data(), nInstances * sizeof(InstanceDev), cudaMemcpyHostToDevice, s));
    DVR_CUDA(cudaStreamSynchronize(s));
    L.ext = ext;
  }
  L.sched = acquireSchedSlot();
  if (!L.sched) {
    setError("dvr_render: could not allocate scheduler scratch");
    return DVR_ERR_CUDA;
  }

  unsigned int *bitmap = nullptr;
  size_t nWords = 0;
  if (stats) {
    DVR_CUDA(cudaMemsetAsync(statsDev, 0, sizeof(DvrRenderStats), s));
    if (nInstances >= 1) {
      const size_t nCells = instances[0].volume->field->nCells;
      nWords = (nCells + 31) / 32;
      DVR_CUDA(scratchAllocAsync((void **)&bitmap, nWords * 4, s));
      DVR_CUDA(cudaMemsetAsync(bitmap, 0, nWords * 4, s));
    }
    L.stats = statsDev;
    L.cellBitmap = bitmap;
  }

  int rc = DVR_OK;
  {
    // the fork / join events are per device: one frame at a time enqueues its sweep (host-side enqueue only)
    static std::mutex sweepMutex;
    std::unique_lock<std::mutex> sweepLock(sweepMutex, std::defer_lock);
    // device-only frames: sweep first (wide, short); mirrored frames: march first, then the thin PCIe-bound sweep
    const bool sweepFirst = L.sweepOutside && !L.fb.outMirror;
    if (L.sweepOutside) {
      sweepLock.lock();
      rc = forkBackgroundSweep(s);
      if (rc == DVR_OK && sweepFirst)
        rc = enqueueBackgroundSweep(L);
    }
    if (rc == DVR_OK && hasSurfaces)
      rc = launchSceneFrame(L, scene, skip && nInstances > 0, s);
    else if (rc == DVR_OK) {
      if (nInstances == 0) {
        // a world without volumes still clears to the background (Raycast_ptx.cu:139-166 with no hit)
        L.nInst = 0;
        rc = launchFrame(L, false, stats, s);
      } else
        rc = launchFrame(L, skip, stats, s);
    }
    if (L.sweepOutside && !sweepFirst && rc == DVR_OK)
      rc = enqueueBackgroundSweep(L);
    if (L.sweepOutside) {
      const int rcj = joinBackgroundSweep(s);
      if (rc == DVR_OK)
        rc = rcj;
    }
  }

  if (stats && bitmap) {
    if (rc == DVR_OK)
      rc = launchPopcount(bitmap, nWords, &statsDev->macrocellsTouched, s);
    cudaFreeAsync(bitmap, s);
  }
  if (L.ext)
    cudaFreeAsync((void *)L.ext, s);
  return rc;
}

int dvr_render(const DvrFrameParams *params, const DvrCamera *camera, const DvrVolumeInstance *instances,
    uint32_t nInstances, const DvrFrameBuffers *buffers, void *stream)
{
  return renderImpl(params, camera, instances, nInstances, buffers, nullptr, false, stream);
}

int dvr_render_instrumented(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instances, uint32_t nInstances, const DvrFrameBuffers *buffers,
    DvrRenderStats *statsDev, void *stream)
{
  return renderImpl(params, camera, instances, nInstances, buffers, statsDev, true, stream);
}

int dvr_render_scene(const DvrFrameParams *params, const DvrCamera *camera, const DvrVolumeInstance *instances,
    uint32_t nInstances, const DvrSceneParams *scene, const DvrFrameBuffers *buffers, void *stream)
{
  if (!scene) {
    setError("dvr_render_scene: null scene parameters");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return renderImpl(params, camera, instances, nInstances, buffers, nullptr, false, stream, scene);
}

// ---- sort-last -------------------------------------------------------------------------------------------

static int fillSync(const DvrPeerSync *in, SyncDev &out)
{
  std::memset(&out, 0, sizeof(out));
  if (!in)
    return DVR_OK;
  if (in->nSignal > (uint32_t)kMaxSlabs || in->nWait > (uint32_t)kMaxSlabs || (in->nWait && !in->wait)) {
    setError("DvrPeerSync: at most 16 signal/wait flags; wait table must not be null");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  out.nSignal = in->nSignal;
  out.signalValue = in->signalValue;
  for (uint32_t i = 0; i < in->nSignal; ++i) {
    if (!in->signal[i]) {
      setError("DvrPeerSync: null signal flag");
      return DVR_ERR_INVALID_ARGUMENT;
    }
    out.signal[i] = in->signal[i];
  }
  out.nWait = in->nWait;
  out.waitValue = in->waitValue;
  out.wait = in->wait;
  out.errorFlag = in->errorFlag;
  return DVR_OK;
}

// kernel parameters of the slab march (shared by dvr_render_partial* and the fused dvr_render_slab_frame)
static void fillPartialLaunch(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    float *partialRgba, float *partialDepth, bool cullToBounds, PartialLaunch &L)
{
  std::memset(&L, 0, sizeof(L));
  L.width = p->width;
  L.height = p->height;
  L.invW = 1.f / (float)p->width;
  L.invH = 1.f / (float)p->height;
  L.integrator = p->integrator;
  L.frameID = p->frameID;
  L.invSamplingRate = p->inverseVolumeSamplingRate;
  L.tilesX = (p->width + kTileW - 1) / kTileW;
  L.tilesY = (p->height + kTileH - 1) / kTileH;
  L.tileX0 = L.tileY0 = 0;
  L.tilesW = L.tilesX;
  L.tilesH = L.tilesY;
  int rect[4];
  if (cullToBounds && screenRectOfBounds(camera, instance, p->width, p->height, rect)) {
    if (rect[2] <= rect[0] || rect[3] <= rect[1]) { // the volume is off screen: nothing to march
      L.tilesW = L.tilesH = 0;
    } else {
      L.tileX0 = (uint32_t)rect[0] / kTileW;
      L.tileY0 = (uint32_t)rect[1] / kTileH;
      L.tilesW = ((uint32_t)rect[2] + kTileW - 1) / kTileW - L.tileX0;
      L.tilesH = ((uint32_t)rect[3] + kTileH - 1) / kTileH - L.tileY0;
    }
  }
  fillCamera(camera, L.cam);
  fillInstance(*instance, L.inst);
  L.partialRgba = (float4 *)partialRgba;
  L.partialDepth = partialDepth;
  L.skip = p->useMacrocellSkipping == DVR_SKIP_ON
      || (p->useMacrocellSkipping == DVR_SKIP_AUTO && instance->volume->emptyFraction >= kSkipAutoThreshold);
}

static int renderPartialImpl(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    float *partialRgba, float *partialDepth, DvrRenderStats *statsDev, void *stream, const DvrPeerSync *sync = nullptr)
{
  if (!p || !camera || !instance || !instance->volume || !partialRgba || !partialDepth) {
    setError("dvr_render_partial: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (!instance->volume->field) {
    setError("dvr_render_partial: instance without a valid volume");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (p->width == 0 || p->height == 0) {
    setError("dvr_render_partial: empty frame");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (p->numIterations != 1 || p->checkerboardID >= 0) {
    setError("dvr_render_partial: needs numIterations == 1 and no checkerboarding");
    return DVR_ERR_UNSUPPORTED;
  }
  if (p->integrator != DVR_INTEGRATOR_RAYCAST && p->integrator != DVR_INTEGRATOR_DEFAULT) {
    setError("dvr_render_partial: only the marching integrators (raycast, default) render partial images");
    return DVR_ERR_UNSUPPORTED;
  }
  if (instance->volume->field->dev.kind != FIELD_STRUCTURED) {
    setError("dvr_render_partial: slab rendering needs a structuredRegular field");
    return DVR_ERR_UNSUPPORTED;
  }
  if (dvr_device_count() <= 0) {
    setError("dvr_render_partial: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  PartialLaunch L;
  fillPartialLaunch(p, camera, instance, partialRgba, partialDepth, p->partialCullToBounds != 0, L);
  {
    const int rcs = fillSync(sync, L.sync);
    if (rcs != DVR_OK)
      return rcs;
  }
  L.sched = acquireSchedSlot();
  if (!L.sched) {
    setError("dvr_render_partial: could not allocate scheduler scratch");
    return DVR_ERR_CUDA;
  }
  unsigned int *bitmap = nullptr;
  size_t nWords = 0;
  if (statsDev) {
    DVR_CUDA(cudaMemsetAsync(statsDev, 0, sizeof(DvrRenderStats), s));
    nWords = (instance->volume->field->nCells + 31) / 32;
    DVR_CUDA(scratchAllocAsync((void **)&bitmap, nWords * 4, s));
    DVR_CUDA(cudaMemsetAsync(bitmap, 0, nWords * 4, s));
    L.stats = statsDev;
    L.cellBitmap = bitmap;
  }
  int rc = launchPartial(L, s);
  if (bitmap) {
    if (rc == DVR_OK)
      rc = launchPopcount(bitmap, nWords, &statsDev->macrocellsTouched, s);
    cudaFreeAsync(bitmap, s);
  }
  return rc;
}

int dvr_render_partial(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    float *partialRgba, float *partialDepth, void *stream)
{
  return renderPartialImpl(p, camera, instance, partialRgba, partialDepth, nullptr, stream);
}

int dvr_render_partial_sync(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    float *partialRgba, float *partialDepth, const DvrPeerSync *sync, void *stream)
{
  return renderPartialImpl(p, camera, instance, partialRgba, partialDepth, nullptr, stream, sync);
}

int dvr_wait_flags(const unsigned int *flags, uint32_t n, uint32_t value, unsigned int *errorFlag, void *stream)
{
  if (!flags || n == 0) {
    setError("dvr_wait_flags: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return launchWaitFlags(flags, n, value, errorFlag, (cudaStream_t)stream);
}

int dvr_render_partial_instrumented(const DvrFrameParams *p, const DvrCamera *camera,
    const DvrVolumeInstance *instance, float *partialRgba, float *partialDepth, DvrRenderStats *statsDev, void *stream)
{
  if (!statsDev) {
    setError("dvr_render_partial_instrumented: statsDev is null");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return renderPartialImpl(p, camera, instance, partialRgba, partialDepth, statsDev, stream);
}

int dvr_composite_over(float *frontRgba, float *frontDepth, const float *backRgba, const float *backDepth,
    size_t pixelBegin, size_t pixelEnd, int backIsInFront, void *stream)
{
  if (!frontRgba || !backRgba) {
    setError("dvr_composite_over: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return launchCompositeOver((float4 *)frontRgba, frontDepth, (const float4 *)backRgba, backDepth, pixelBegin,
      pixelEnd, backIsInFront != 0, (cudaStream_t)stream);
}

int dvr_resolve(const DvrFrameParams *p, const float *partialRgba, const float *partialDepth, uint32_t objId,
    uint32_t instId, const DvrFrameBuffers *b, size_t pixelBegin, size_t pixelEnd, void *stream)
{
  if (!p || !partialRgba || !b || !b->colorAccumulation || !b->outColor) {
    setError("dvr_resolve: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  ResolveLaunch R;
  std::memset(&R, 0, sizeof(R));
  R.width = p->width;
  R.height = p->height;
  R.format = p->format;
  R.frameID = p->frameID;
  R.background = make_float4(p->background[0], p->background[1], p->background[2], p->background[3]);
  R.bgTex = p->backgroundImage ? p->backgroundImage->tex : 0;
  R.invW = 1.f / (float)p->width;
  R.invH = 1.f / (float)p->height;
  R.centered = p->integrator == DVR_INTEGRATOR_RAYCAST ? 1 : 0;
  R.fb.accum = (float4 *)b->colorAccumulation;
  R.fb.outU32 = (uint32_t *)b->outColor;
  R.fb.outF32 = (float4 *)b->outColor;
  R.fb.outMirror = b->outColorMirror;
  R.fb.depth = b->depth;
  R.fb.depthMirror = b->depth ? b->depthMirror : nullptr;
  R.fb.primId = b->primId;
  R.fb.objId = b->objId;
  R.fb.instId = b->instId;
  R.fb.albedo = b->albedo;
  R.fb.normal = b->normal;
  R.partialRgba = (const float4 *)partialRgba;
  R.partialDepth = partialDepth;
  R.objId = objId;
  R.instId = instId;
  R.pixelBegin = pixelBegin;
  R.pixelEnd = pixelEnd > (size_t)p->width * p->height ? (size_t)p->width * p->height : pixelEnd;
  return launchResolve(R, (cudaStream_t)stream);
}

static int fillPeerResolveLaunch(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    const float *const *partialRgba, const float *const *partialDepth, uint32_t nSlabs, uint32_t objId,
    uint32_t instId, const DvrFrameBuffers *b, size_t pixelBegin, size_t pixelEnd, PeerResolveLaunch &L)
{
  std::memset(&L, 0, sizeof(L));
  ResolveLaunch &R = L.r;
  R.width = p->width;
  R.height = p->height;
  R.format = p->format;
  R.frameID = p->frameID;
  R.background = make_float4(p->background[0], p->background[1], p->background[2], p->background[3]);
  R.bgTex = p->backgroundImage ? p->backgroundImage->tex : 0;
  R.invW = 1.f / (float)p->width;
  R.invH = 1.f / (float)p->height;
  R.centered = p->integrator == DVR_INTEGRATOR_RAYCAST ? 1 : 0;
  R.fb.accum = (float4 *)b->colorAccumulation;
  R.fb.outU32 = (uint32_t *)b->outColor;
  R.fb.outF32 = (float4 *)b->outColor;
  R.fb.outMirror = b->outColorMirror;
  R.fb.depth = b->depth;
  R.fb.depthMirror = b->depth ? b->depthMirror : nullptr;
  R.fb.primId = b->primId;
  R.fb.objId = b->objId;
  R.fb.instId = b->instId;
  R.fb.albedo = b->albedo;
  R.fb.normal = b->normal;
  R.objId = objId;
  R.instId = instId;
  R.pixelBegin = pixelBegin;
  R.pixelEnd = pixelEnd > (size_t)p->width * p->height ? (size_t)p->width * p->height : pixelEnd;
  fillCamera(camera, L.cam);
  L.invW = 1.f / (float)p->width;
  L.invH = 1.f / (float)p->height;
  L.nSlabs = (int)nSlabs;
  for (uint32_t i = 0; i < nSlabs; ++i) {
    if (!partialRgba[i]) {
      setError("dvr_composite_resolve_peers: null partial image");
      return DVR_ERR_INVALID_ARGUMENT;
    }
    L.rgba[i] = (const float4 *)partialRgba[i];
    L.depth[i] = partialDepth ? partialDepth[i] : nullptr;
  }
  L.integrator = p->integrator;
  L.identity = 1u;
  if (instance && instance->volume && instance->volume->field) {
    InstanceDev tmp;
    fillInstance(*instance, tmp);
    L.cull = 1;
    L.boundsLo = tmp.v.f.boundsLo;
    L.boundsHi = tmp.v.f.boundsHi;
    std::memcpy(L.xfm, tmp.xfm, sizeof(L.xfm));
    L.identity = tmp.identity;
    int rect[4];
    if (screenRectOfBounds(camera, instance, p->width, p->height, rect)) {
      L.missValid = 1;
      L.missX0 = rect[0];
      L.missY0 = rect[1];
      L.missX1 = rect[2];
      L.missY1 = rect[3];
    }
  }
  return DVR_OK;
}

static int compositeResolveImpl(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    const float *const *partialRgba, const float *const *partialDepth, uint32_t nSlabs, uint32_t objId,
    uint32_t instId, const DvrFrameBuffers *b, size_t pixelBegin, size_t pixelEnd, const DvrPeerSync *sync,
    void *stream)
{
  if (!p || !camera || !partialRgba || !b || !b->colorAccumulation || !b->outColor || nSlabs == 0) {
    setError("dvr_composite_resolve_peers: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (nSlabs > (uint32_t)kMaxSlabs) {
    setError("dvr_composite_resolve_peers: at most 16 slabs");
    return DVR_ERR_UNSUPPORTED;
  }
  PeerResolveLaunch L;
  {
    const int rcf = fillPeerResolveLaunch(p, camera, instance, partialRgba, partialDepth, nSlabs, objId, instId, b,
        pixelBegin, pixelEnd, L);
    if (rcf != DVR_OK)
      return rcf;
  }
  const int rcs = fillSync(sync, L.sync);
  if (rcs != DVR_OK)
    return rcs;
  return launchPeerResolve(L, (cudaStream_t)stream);
}

int dvr_composite_resolve_peers(const DvrFrameParams *p, const DvrCamera *camera, const float *const *partialRgba,
    const float *const *partialDepth, uint32_t nSlabs, uint32_t objId, uint32_t instId, const DvrFrameBuffers *b,
    size_t pixelBegin, size_t pixelEnd, void *stream)
{
  return compositeResolveImpl(p, camera, nullptr, partialRgba, partialDepth, nSlabs, objId, instId, b, pixelBegin,
      pixelEnd, nullptr, stream);
}

int dvr_composite_resolve_peers_sync(const DvrFrameParams *p, const DvrCamera *camera,
    const DvrVolumeInstance *instance, const float *const *partialRgba, const float *const *partialDepth,
    uint32_t nSlabs, uint32_t objId, uint32_t instId, const DvrFrameBuffers *b, size_t pixelBegin, size_t pixelEnd,
    const DvrPeerSync *sync, void *stream)
{
  return compositeResolveImpl(p, camera, instance, partialRgba, partialDepth, nSlabs, objId, instId, b, pixelBegin,
      pixelEnd, sync, stream);
}

int dvr_render_slab_frame(const DvrFrameParams *p, const DvrCamera *camera, const DvrVolumeInstance *instance,
    uint32_t objId, uint32_t instId, const DvrFrameBuffers *b, const DvrSlabExchange *x, void *stream)
{
  if (!p || !camera || !instance || !instance->volume || !instance->volume->field || !b || !b->colorAccumulation
      || !b->outColor || !x || !x->partialRgba || !x->partialDepth || !x->regionFlags || !x->resolvedFlags
      || !x->regionDone) {
    setError("dvr_render_slab_frame: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (x->nRanks == 0 || x->nRanks > (uint32_t)kMaxSlabs || x->rank >= x->nRanks || x->seq == 0 || x->maxRegions == 0) {
    setError("dvr_render_slab_frame: 1..16 ranks, rank < nRanks, seq > 0, maxRegions > 0");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (p->width == 0 || p->height == 0) {
    setError("dvr_render_slab_frame: empty frame");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (p->numIterations != 1 || p->checkerboardID >= 0
      || (p->integrator != DVR_INTEGRATOR_RAYCAST && p->integrator != DVR_INTEGRATOR_DEFAULT)
      || instance->volume->field->dev.kind != FIELD_STRUCTURED) {
    setError("dvr_render_slab_frame: marching integrators, 1 sample per pixel, no checkerboarding, structuredRegular slab");
    return DVR_ERR_UNSUPPORTED;
  }
  if (dvr_device_count() <= 0) {
    setError("dvr_render_slab_frame: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  for (uint32_t i = 0; i < x->nRanks; ++i)
    if (!x->partialRgba[i] || !x->partialDepth[i] || !x->regionFlags[i] || !x->resolvedFlags[i]) {
      setError("dvr_render_slab_frame: null per-rank pointer");
      return DVR_ERR_INVALID_ARGUMENT;
    }
  SlabFrameLaunch S;
  std::memset(&S, 0, sizeof(S));
  fillPartialLaunch(p, camera, instance, const_cast<float *>(x->partialRgba[x->rank]),
      const_cast<float *>(x->partialDepth[x->rank]), true, S.m);
  const size_t npx = (size_t)p->width * p->height;
  if (npx >= 0xffffffffull) { // the background strip indexes its pixels with 32 bits
    setError("dvr_render_slab_frame: frames of 2^32 pixels or more are not supported");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  {
    const int rcf = fillPeerResolveLaunch(p, camera, instance, x->partialRgba, x->partialDepth, x->nRanks, objId, instId,
        b, 0, npx, S.c);
    if (rcf != DVR_OK)
      return rcf;
  }
  S.c.sync.errorFlag = x->errorFlag;
  S.nRanks = x->nRanks;
  S.rank = x->rank;
  S.seq = x->seq;
  const uint32_t nTiles = S.m.tilesW * S.m.tilesH;
  S.tilesPerRegion = std::max<uint32_t>(32u, (nTiles + x->maxRegions - 1) / x->maxRegions);
  S.nRegions = (nTiles + S.tilesPerRegion - 1) / S.tilesPerRegion;
  S.regionDone = x->regionDone;
  for (uint32_t i = 0; i < x->nRanks; ++i) {
    S.regionFlags[i] = x->regionFlags[i];
    S.resolvedFlags[i] = x->resolvedFlags[i];
  }
  S.myRegionFlags = x->regionFlags[x->rank];
  S.myResolved = x->resolvedFlags[x->rank];
  S.waitAllResolved = x->waitAllResolved;
  S.timing = x->timing;
  {
    static const unsigned spinNs = []() {
      const char *e = std::getenv("DVR_B200_SPIN_NS");
      const int v = e ? std::atoi(e) : 0;
      return v > 0 ? (unsigned)v : 400u; // 100 / 400..3200 / 1000 / 2000 ns at N = 2: 1881 / 1917 / 1917 / 1912 frames/s (call X)
    }();
    S.spinSleepNs = spinNs;
    static const unsigned capNs = []() {
      const char *e = std::getenv("DVR_B200_SPIN_CAP_NS");
      const int v = e ? std::atoi(e) : 0;
      return v > 0 ? (unsigned)v : 3200u;
    }();
    S.spinSleepCapNs = capNs > spinNs ? capNs : spinNs;
    static const unsigned dbg = []() {
      const char *e = std::getenv("DVR_B200_SLAB_DEBUG");
      return e ? (unsigned)std::atoi(e) : 0u;
    }();
    S.debugFlags = dbg;
  }
  { // this rank's share of the pixels outside the window: contiguous strips, 256-pixel aligned
    size_t per = (npx + x->nRanks - 1) / x->nRanks;
    per = (per + 255) / 256 * 256;
    S.bgPixelBegin = std::min(npx, per * x->rank);
    S.bgPixelEnd = std::min(npx, S.bgPixelBegin + per);
  }
  S.m.sched = acquireSchedSlot();
  if (!S.m.sched) {
    setError("dvr_render_slab_frame: could not allocate scheduler scratch");
    return DVR_ERR_CUDA;
  }
  return launchSlabFrame(S, (cudaStream_t)stream);
}

int dvr_ipc_alloc(size_t bytes, void **devPtr, unsigned char handle[DVR_IPC_HANDLE_BYTES])
{
  static_assert(sizeof(cudaIpcMemHandle_t) == DVR_IPC_HANDLE_BYTES, "IPC handle size");
  if (!devPtr || !handle || bytes == 0) {
    setError("dvr_ipc_alloc: bad argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  DVR_CUDA(cudaMalloc(devPtr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *devPtr);
  if (e != cudaSuccess) {
    cudaFree(*devPtr);
    *devPtr = nullptr;
    return cudaFail(e, "cudaIpcGetMemHandle");
  }
  std::memcpy(handle, &h, sizeof(h));
  return DVR_OK;
}

int dvr_ipc_open(const unsigned char handle[DVR_IPC_HANDLE_BYTES], void **devPtr)
{
  if (!devPtr || !handle) {
    setError("dvr_ipc_open: bad argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  DVR_CUDA(cudaIpcOpenMemHandle(devPtr, h, cudaIpcMemLazyEnablePeerAccess));
  return DVR_OK;
}

int dvr_ipc_close(void *devPtr)
{
  if (devPtr)
    DVR_CUDA(cudaIpcCloseMemHandle(devPtr));
  return DVR_OK;
}

int dvr_ipc_free(void *devPtr)
{
  if (devPtr)
    DVR_CUDA(cudaFree(devPtr));
  return DVR_OK;
}

int dvr_bounds_screen_rect(const DvrCamera *camera, const float boundsLo[3], const float boundsHi[3], uint32_t width,
    uint32_t height, int32_t rect[4])
{
  if (!camera || !boundsLo || !boundsHi || !rect || width == 0 || height == 0) {
    setError("dvr_bounds_screen_rect: invalid argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  int r[4] = {0, 0, (int)width, (int)height};
  const bool ok = screenRectOfBox(camera, make_float3(boundsLo[0], boundsLo[1], boundsLo[2]),
      make_float3(boundsHi[0], boundsHi[1], boundsHi[2]), width, height, r);
  for (int i = 0; i < 4; ++i)
    rect[i] = r[i];
  return ok ? 1 : 0;
}

int dvr_selftest_lattice_advance(uint32_t count, uint64_t seed, uint32_t *mismatchesOut, void *stream)
{
  if (!mismatchesOut) {
    setError("dvr_selftest_lattice_advance: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (dvr_device_count() <= 0) {
    setError("dvr_selftest_lattice_advance: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  unsigned int *d = nullptr;
  DVR_CUDA(cudaMalloc(&d, sizeof(unsigned int)));
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(d, 0, sizeof(unsigned int), s);
  int rc = count ? launchSelftestLattice(count, seed, d, s) : DVR_OK;
  if (rc == DVR_OK) {
    const cudaError_t e = cudaMemcpyAsync(mismatchesOut, d, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e != cudaSuccess || e2 != cudaSuccess)
      rc = cudaFail(e != cudaSuccess ? e : e2, "dvr_selftest_lattice_advance");
  }
  cudaFree(d);
  return rc;
}

int dvr_scale_vec3(const float *accumVec3, float *outVec3, size_t nPixels, float scale, void *stream)
{
  if (!accumVec3 || !outVec3) {
    setError("dvr_scale_vec3: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  return launchScaleVec3(accumVec3, outVec3, nPixels, scale, (cudaStream_t)stream);
}

} // extern "C"
