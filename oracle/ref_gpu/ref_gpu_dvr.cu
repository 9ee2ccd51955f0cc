// ref_gpu_dvr.cu — O-gpu: the reference's OWN device code for the DVR path, compiled for sm_100a.
// TEST INFRASTRUCTURE (oracle).  Built by oracle/Makefile into oracle/_ref/libref_gpu_dvr.so from the
// reference headers where they lie under /root/reference (nothing is copied into this repository).
//
// What runs unmodified from the reference (devices/rtx/gpu/*.h): createScreenSample, makePrimaryRay,
// cameraCreateRay, rayMarchAllVolumes, detail::rayMarchVolume, _rayMarchVolume,
// SpatialFieldSampler<cudaTextureObject_t>, classifySample, getBackground, accumulateValue,
// accumResults (tonemap / inverseTonemap / writeOutputColor) — with hardware tex3D / tex1D and
// cuRAND Philox, exactly as the OptiX raygen programs call them.
// What is restated here because it lives in OptiX programs or host classes that cannot be built
// without OptiX / ANARI-SDK:
//   - the raygen "no surface hit" branch            renderer/Raycast_ptx.cu:60-179, DirectLight_ptx.cu:294-418
//   - the volume AABB search (RT traversal + IS/CH) scene/Intersectors_ptx.cu:248-274, gpu/populateHit.h:370-390
//   - host texture set-up                           StructuredRegularField.cpp:98-194, TransferFunction1D.cpp:152-186
//   - frame reset fills                             frame/Frame.cu:590-660
//   - the background image's RGBA8 staging pass     utility/CudaImageTexture.cpp:43-58,84-101,139-226,316-345
//     (that file needs helium's Array class; its per-component conversions are restated with the same glm call)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gpu/gpu_util.h"
#include "gpu/intersectRay.h"
#include "gpu/createScreenSample.h"
#include "gpu/volumeIntegration.h"

// the reference's own majorant-grid build (host class + kernels), compiled from where it lies
#include "scene/volume/space_skipping/UniformGrid.cu"

#include <glm/gtc/color_space.hpp>

#include "dvr_b200.h" // POD parameter structs only (DvrFrameParams, DvrCamera, DvrFrameBuffers)

using namespace visrtx;

struct RefInstanceXfm
{
  float m[12]; // world -> object
  int identity;
};

__constant__ FrameGPUData frameData; // same launch-parameter symbol name as the reference (Renderer.cpp:733)
__device__ const RefInstanceXfm *g_xfms;

// Volume-BVH trace replacement: closest clamped AABB entry among the volume instances, skipping the
// one hit last (Intersectors_ptx.cu:250-252), filling VolumeHit like populateVolumeHit.
__device__ void refgpu_shim_trace(unsigned long long, float3 org, float3 dir, float tmin, float tmax, unsigned ssHi,
    unsigned ssLo, unsigned dataHi, unsigned dataLo, unsigned bvhSelection)
{
  if (bvhSelection != 0u)
    return; // surfaces: a volume-only world has none => miss
  ScreenSample &ss = *(ScreenSample *)detail::unpackPointer(ssHi, ssLo);
  VolumeHit &hit = *(VolumeHit *)detail::unpackPointer(dataHi, dataLo);
  const FrameGPUData &fd = *ss.frameData;
  int best = -1;
  box1 bt;
  vec3 bo, bd;
  for (int i = 0; i < (int)fd.world.numVolumeInstances; ++i) {
    const uint32_t objID = 0u, instID = (uint32_t)i;
    if (hit.lastVolID == objID && hit.lastInstID == instID)
      continue;
    const auto &inst = fd.world.volumeInstances[i];
    const VolumeGPUData &vd = fd.registry.volumes[inst.volumes[0]];
    vec3 lo(org.x, org.y, org.z), ld(dir.x, dir.y, dir.z);
    const RefInstanceXfm &x = g_xfms[i];
    if (!x.identity) {
      const float *m = x.m;
      const vec3 o = lo, d = ld;
      lo = vec3(m[0] * o.x + m[1] * o.y + m[2] * o.z + m[3], m[4] * o.x + m[5] * o.y + m[6] * o.z + m[7],
          m[8] * o.x + m[9] * o.y + m[10] * o.z + m[11]);
      ld = vec3(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z,
          m[8] * d.x + m[9] * d.y + m[10] * d.z);
    }
    const auto &bounds = vd.bounds;
    const vec3 mins = (bounds.lower - lo) * (1.f / ld);
    const vec3 maxs = (bounds.upper - lo) * (1.f / ld);
    const vec3 nears = glm::min(mins, maxs);
    const vec3 fars = glm::max(mins, maxs);
    box1 t(glm::compMax(nears), glm::compMin(fars));
    if (!(t.lower < t.upper))
      continue;
    if (t.upper < tmin || t.lower > tmax)
      continue; // the traversal never visits an AABB outside the ray interval
    const box1 rayt{tmin, tmax};
    t.lower = clamp(t.lower, rayt);
    t.upper = clamp(t.upper, rayt);
    if (best < 0 || t.lower < bt.lower) {
      best = i;
      bt = t;
      bo = lo;
      bd = ld;
    }
  }
  if (best < 0)
    return;
  const auto &inst = fd.world.volumeInstances[best];
  hit.foundHit = true;
  hit.volume = &fd.registry.volumes[inst.volumes[0]];
  hit.instance = &inst;
  hit.lastVolID = 0u;
  hit.lastInstID = (uint32_t)best;
  hit.localRay.org = bo;
  hit.localRay.dir = bd;
  hit.localRay.t.lower = bt.lower;
  hit.localRay.t.upper = bt.upper;
}

// raygen, volume-only world.  centerPixel=true/1 iteration == raycast; otherwise default/directLight.
__global__ void refgpu_raygen(int centerPixel)
{
  auto &rendererParams = frameData.renderer;
  auto ss = createScreenSample(frameData);
  if (pixelOutOfFrame(ss.pixel, frameData.fb))
    return;
  const int iters = centerPixel ? 1 : frameData.renderer.numIterations;
  for (int i = 0; i < iters; i++) {
    auto ray = makePrimaryRay(ss, centerPixel != 0);
    vec3 outputColor(0.f);
    vec3 outputNormal = ray.dir;
    float outputOpacity = 0.f;
    float depth = 1e30f;
    uint32_t primID = ~0u, objID = ~0u, instID = ~0u;

    vec3 color(0.f);
    float opacity = 0.f;
    uint32_t vObjID = ~0u, vInstID = ~0u;
    const float volumeDepth = rayMarchAllVolumes(ss, ray, 0 /*RayType::PRIMARY*/, ray.t.upper,
        rendererParams.inverseVolumeSamplingRate, color, opacity, vObjID, vInstID);
    depth = min(depth, volumeDepth);
    primID = 0;
    objID = vObjID;
    instID = vInstID;
    color *= opacity;
    const auto bg = getBackground(frameData, ss.screen, ray.dir);
    accumulateValue(color, vec3(bg), opacity);
    accumulateValue(opacity, bg.w, opacity);
    accumulateValue(outputColor, color, outputOpacity);
    accumulateValue(outputOpacity, opacity, outputOpacity);

    accumResults(frameData.fb, ss.pixel, vec4(outputColor, outputOpacity), depth, outputColor, outputNormal, primID,
        objID, instID, i);
  }
}

// raygen of the dpt renderer for a volume-only world: renderer/DiffusePathTracer_ptx.cu:82-215 with the
// surface branch (intersectSurface always misses here) removed.  sampleDistanceAllVolumes, _sampleDistance,
// dda3, sampleUnitSphere and accumResults are the reference's own.
struct RefPathData // DiffusePathTracer_ptx.cu:45-50 without the surface Hit
{
  int depth{0};
  vec3 Lw{1.f};
};

__global__ void refgpu_raygen_dpt()
{
  auto &rendererParams = frameData.renderer;
  auto &dptParams = rendererParams.params.dpt;
  RefPathData pathData;
  auto ss = createScreenSample(frameData);
  if (pixelOutOfFrame(ss.pixel, frameData.fb))
    return;
  for (int i = 0; i < frameData.renderer.numIterations; i++) {
    auto ray = makePrimaryRay(ss);
    auto tmax = ray.t.upper;
    const auto bg = getBackground(frameData, ss.screen, ray.dir);
    vec3 outColor(bg);
    vec3 outNormal = ray.dir;
    float outDepth = tmax;
    uint32_t primID = ~0u, objID = ~0u, instID = ~0u;
    while (true) {
      float volumeOpacity = 0.f;
      vec3 volumeColor(0.f);
      float Tr = 0.f;
      uint32_t vObjID = ~0u, vInstID = ~0u;
      const float volumeDepth = sampleDistanceAllVolumes(
          ss, ray, 0 /*RayType::DIFFUSE_RADIANCE*/, ray.t.upper, volumeColor, volumeOpacity, Tr, vObjID, vInstID);
      const bool volumeHit = Tr < 1.f;
      if (!volumeHit)
        break;
      if (pathData.depth++ >= dptParams.maxDepth) {
        pathData.Lw = vec3(0.f);
        break;
      }
      const vec3 pos = ray.org + volumeDepth * ray.dir;
      pathData.Lw *= volumeColor;
      float P = glm::compMax(pathData.Lw);
      if (P < .2f) {
        if (curand_uniform(&ss.rs) > P) {
          pathData.Lw = vec3(0.f);
          break;
        }
        pathData.Lw /= P;
      }
      const vec3 scatterDir = sampleUnitSphere(ss.rs, -ray.dir);
      ray.org = pos;
      ray.dir = scatterDir;
      ray.t.lower = 0.f;
      ray.t.upper = rendererParams.occlusionDistance;
      // the reference's `if (pathData.depth == 0)` block can never run after the increment above
    }
    vec3 Ld(rendererParams.ambientIntensity);
    vec3 color = pathData.depth ? pathData.Lw * Ld : vec3(bg);
    accumResults(frameData.fb, ss.pixel, vec4(color, 1.f), outDepth, outColor, outNormal, primID, objID, instID, i);
  }
}

// raygen of the `test` renderer, renderer/Test_ptx.cu:52-69 verbatim
__global__ void refgpu_raygen_test()
{
  auto ss = createScreenSample(frameData);
  if (pixelOutOfFrame(ss.pixel, frameData.fb))
    return;
  auto ray = makePrimaryRay(ss);
  accumResults(frameData.fb, ss.pixel, vec4(ray.dir, 1.f), 1.f, ray.dir, -ray.dir, ~0u, ~0u, ~0u);
}

template <typename T>
__global__ void refgpu_fill(T *p, size_t n, T v)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}
template <typename T>
static void fill(T *p, size_t n, T v, cudaStream_t s)
{
  if (p && n)
    refgpu_fill<T><<<1184, 256, 0, s>>>(p, n, v);
}

struct RefField
{
  cudaArray_t arr = nullptr;
  cudaTextureObject_t tex = 0;
  void *nvdb = nullptr;
  SpatialFieldGPUData gpu{};
  box3 bounds;
  float stepSize = 0.f;
  ivec3 dims{0}; // what the field passes to UniformGrid::init (voxel counts / NanoVDB index-bbox extent)
};
struct RefVolume
{
  RefField *field = nullptr;
  cudaArray_t arr = nullptr;
  cudaTextureObject_t tex = 0;
  VolumeGPUData gpu{};
  ivec3 gridDims{0};            // injected delta-tracking grid (refgpu_volume_set_grid)
  float *maxOpacities = nullptr;
};

static thread_local char g_err[512];
#define RCK(x)                                                                                      \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) {                                                                        \
      snprintf(g_err, sizeof(g_err), "%s: %s", #x, cudaGetErrorString(e_));                         \
      return -3;                                                                                    \
    }                                                                                               \
  } while (0)

extern "C" {

const char *refgpu_last_error() { return g_err; }

// StructuredRegularField::finalize + gpuData (f32 / u8 / u16 host data)
int refgpu_field_create(const void *hostVoxels, int dataType, const uint32_t dims[3], const float origin[3],
    const float spacing[3], int nearest, RefField **out)
{
  auto *f = new RefField();
  int bits = 32;
  cudaChannelFormatKind kind = cudaChannelFormatKindFloat;
  if (dataType == DVR_UFIXED8) { bits = 8; kind = cudaChannelFormatKindUnsigned; }
  else if (dataType == DVR_FIXED8) { bits = 8; kind = cudaChannelFormatKindSigned; }
  else if (dataType == DVR_UFIXED16) { bits = 16; kind = cudaChannelFormatKindUnsigned; }
  else if (dataType == DVR_FIXED16) { bits = 16; kind = cudaChannelFormatKindSigned; }
  else if (dataType != DVR_FLOAT32) { snprintf(g_err, sizeof(g_err), "unsupported type"); return -1; }
  auto desc = cudaCreateChannelDesc(bits, 0, 0, 0, kind);
  RCK(cudaMalloc3DArray(&f->arr, &desc, make_cudaExtent(dims[0], dims[1], dims[2])));
  cudaMemcpy3DParms cp;
  std::memset(&cp, 0, sizeof(cp));
  cp.srcPtr = make_cudaPitchedPtr(const_cast<void *>(hostVoxels), dims[0] * (bits / 8), dims[0], dims[1]);
  cp.dstArray = f->arr;
  cp.extent = make_cudaExtent(dims[0], dims[1], dims[2]);
  cp.kind = cudaMemcpyDefault;
  RCK(cudaMemcpy3D(&cp));
  cudaResourceDesc rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = f->arr;
  cudaTextureDesc td;
  std::memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = nearest ? cudaFilterModePoint : cudaFilterModeLinear;
  td.readMode = kind == cudaChannelFormatKindFloat ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
  td.normalizedCoords = 1;
  RCK(cudaCreateTextureObject(&f->tex, &rd, &td, nullptr));
  const vec3 o(origin[0], origin[1], origin[2]), sp(spacing[0], spacing[1], spacing[2]);
  f->gpu.type = SpatialFieldType::STRUCTURED_REGULAR;
  f->gpu.data.structuredRegular.texObj = f->tex;
  f->gpu.data.structuredRegular.origin = o;
  f->gpu.data.structuredRegular.spacing = sp;
  f->gpu.data.structuredRegular.invSpacing = vec3(1.f) / (sp * vec3(dims[0], dims[1], dims[2]));
  f->gpu.grid = UniformGridData{};
  f->bounds = box3(o, o + ((vec3(dims[0], dims[1], dims[2]) - 1.f) * sp));
  f->stepSize = glm::compMin(sp / 2.f);
  f->dims = ivec3(dims[0], dims[1], dims[2]);
  *out = f;
  return 0;
}

// NvdbRegularField::finalize + gpuData (spatial_field/NvdbRegularField.cpp:64-143): one serialized grid
int refgpu_field_create_nvdb(const void *hostBlob, size_t bytes, RefField **out)
{
  auto *f = new RefField();
  const auto *meta = reinterpret_cast<const nanovdb::GridData *>(hostBlob);
  if (!meta->isValid() || meta->mGridCount != 1) {
    snprintf(g_err, sizeof(g_err), "invalid NanoVDB buffer");
    return -1;
  }
  RCK(cudaMalloc(&f->nvdb, bytes));
  RCK(cudaMemcpy(f->nvdb, hostBlob, bytes, cudaMemcpyHostToDevice));
  const auto &wb = meta->mWorldBBox;
  f->bounds = box3(vec3(wb.min()[0], wb.min()[1], wb.min()[2]), vec3(wb.max()[0], wb.max()[1], wb.max()[2]));
  const vec3 voxelSize(meta->mVoxelSize[0], meta->mVoxelSize[1], meta->mVoxelSize[2]);
  f->stepSize = glm::compMin(voxelSize) / 2.0f;
  f->gpu.type = SpatialFieldType::NANOVDB_REGULAR;
  f->gpu.data.nvdbRegular.voxelSize = voxelSize;
  f->gpu.data.nvdbRegular.origin = f->bounds.lower;
  f->gpu.data.nvdbRegular.gridData = f->nvdb;
  f->gpu.data.nvdbRegular.gridType = meta->mGridType;
  f->gpu.grid = UniformGridData{};
  {
    const auto gridSize = meta->indexBBox().dim(); // NvdbRegularField.cpp:152-153
    f->dims = ivec3(gridSize[0], gridSize[1], gridSize[2]);
  }
  *out = f;
  return 0;
}

int refgpu_field_destroy(RefField *f)
{
  if (!f) return 0;
  if (f->tex) cudaDestroyTextureObject(f->tex);
  if (f->arr) cudaFreeArray(f->arr);
  if (f->nvdb) cudaFree(f->nvdb);
  delete f;
  return 0;
}

// TransferFunction1D::createTFTexture + gpuData
int refgpu_volume_create(RefField *field, const float *tfRgba, const float valueRange[2], float unitDistance,
    uint32_t id, RefVolume **out)
{
  auto *v = new RefVolume();
  v->field = field;
  auto desc = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
  RCK(cudaMallocArray(&v->arr, &desc, DVR_TF_SIZE));
  RCK(cudaMemcpy2DToArray(v->arr, 0, 0, tfRgba, DVR_TF_SIZE * 16, DVR_TF_SIZE * 16, 1, cudaMemcpyHostToDevice));
  cudaResourceDesc rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = v->arr;
  cudaTextureDesc td;
  std::memset(&td, 0, sizeof(td));
  td.addressMode[0] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 1;
  RCK(cudaCreateTextureObject(&v->tex, &rd, &td, nullptr));
  v->gpu = VolumeGPUData{};
  v->gpu.id = id;
  v->gpu.type = VolumeType::TF1D;
  v->gpu.bounds = field->bounds;
  v->gpu.stepSize = field->stepSize;
  v->gpu.data.tf1d.tfTex = v->tex;
  v->gpu.data.tf1d.valueRange = box1(valueRange[0], valueRange[1]);
  v->gpu.data.tf1d.oneOverUnitDistance = 1.0f / unitDistance;
  v->gpu.data.tf1d.field = 0; // patched per launch
  v->gpu.data.tf1d.uniformColor = vec3(1.f);
  v->gpu.data.tf1d.uniformOpacity = 1.f;
  *out = v;
  return 0;
}

// The delta-tracking grid of the volume's field (UniformGridData).  The reference builds it in
// UniformGrid.cu with two defects (non-conservative cell ranges, majorants from the wrong value
// range) that make its tracker biased; the checker therefore walks the SAME reference tracker code over a grid
// handed in by the test (the product's own grid), which isolates the tracker from the grid build.
int refgpu_volume_set_grid(RefVolume *v, const int dims[3], const float *hostMaxOpacities)
{
  const size_t n = (size_t)dims[0] * dims[1] * dims[2];
  if (v->maxOpacities) cudaFree(v->maxOpacities);
  RCK(cudaMalloc(&v->maxOpacities, n * sizeof(float)));
  RCK(cudaMemcpy(v->maxOpacities, hostMaxOpacities, n * sizeof(float), cudaMemcpyDefault)); // host or device source
  v->gridDims = ivec3(dims[0], dims[1], dims[2]);
  return 0;
}

// The reference's OWN grid, defects included: StructuredRegularField::buildGrid / NvdbRegularField::buildGrid
// (init + buildGrid on the field's gpuData) followed by TransferFunction1D::finalize's
// computeMaxOpacities(stream, tfTex, tfDim) with the default {0,1} range (TransferFunction1D.cpp:74-77).
// Optionally copies the majorants to the host.
int refgpu_volume_build_reference_grid(RefVolume *v, int dimsOut[3], float *hostOut, size_t capacity)
{
  UniformGrid g;
  g.init(v->field->dims, v->field->bounds);
  g.buildGrid(v->field->gpu);
  RCK(cudaDeviceSynchronize());
  g.computeMaxOpacities(nullptr, v->tex, DVR_TF_SIZE);
  RCK(cudaDeviceSynchronize());
  RCK(cudaGetLastError());
  if (v->maxOpacities) cudaFree(v->maxOpacities);
  v->maxOpacities = g.m_maxOpacities; // ownership moves to the volume
  v->gridDims = g.m_dims;
  cudaFree(g.m_valueRanges);
  const size_t n = (size_t)g.m_dims.x * g.m_dims.y * g.m_dims.z;
  if (dimsOut) {
    dimsOut[0] = g.m_dims.x;
    dimsOut[1] = g.m_dims.y;
    dimsOut[2] = g.m_dims.z;
  }
  if (hostOut) {
    if (capacity < n) {
      snprintf(g_err, sizeof(g_err), "capacity %zu < %zu cells", capacity, n);
      return -2;
    }
    RCK(cudaMemcpy(hostOut, v->maxOpacities, n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int refgpu_volume_destroy(RefVolume *v)
{
  if (!v) return 0;
  if (v->maxOpacities) cudaFree(v->maxOpacities);
  if (v->tex) cudaDestroyTextureObject(v->tex);
  if (v->arr) cudaFreeArray(v->arr);
  delete v;
  return 0;
}

struct RefInstance
{
  RefVolume *volume;
  float worldToObject[12];
  uint32_t instanceId;
  uint32_t _pad;
};

// persistent per-scene device tables so that repeated launches (bench) cost one constant upload,
// like Frame::upload() of the reference (Frame.cu:278)
struct RefScene
{
  SpatialFieldGPUData *fields = nullptr;
  VolumeGPUData *volumes = nullptr;
  InstanceVolumeGPUData *instances = nullptr;
  DeviceObjectIndex *volIdx = nullptr;
  RefInstanceXfm *xfms = nullptr;
  CameraGPUData *camera = nullptr;
  int n = 0;
  bool hasGrid = true;
  cudaTextureObject_t bgTex = 0; // Renderer::m_backgroundTexture (0: BackgroundMode::COLOR)
};

// Renderer::finalize (Renderer.cpp:172-179): acquireCUDAArrayUint8 + makeCudaTextureObject(array, true, "linear")
struct RefImage
{
  cudaArray_t arr = nullptr;
  cudaTextureObject_t tex = 0;
};

int refgpu_image_create(const void *pixels, int componentType, int channels, uint32_t w, uint32_t h, RefImage **out)
{
  const size_t n = (size_t)w * h;
  int nc = channels;
  std::vector<uint8_t> staging(n * 4);
  size_t o = 0;
  for (size_t i = 0; i < n * (size_t)channels; ++i) { // transformToStagingBufferUint8 + convertComponentUint8
    uint8_t c;
    if (componentType == DVR_IMAGE_FLOAT32)
      c = uint8_t(((const float *)pixels)[i] * 255);
    else if (componentType == DVR_IMAGE_UFIXED16) {
      constexpr auto maxVal = float(std::numeric_limits<uint16_t>::max());
      c = uint8_t((((const uint16_t *)pixels)[i] / maxVal) * 255);
    } else if (componentType == DVR_IMAGE_UFIXED32) {
      constexpr auto maxVal = float(std::numeric_limits<uint32_t>::max());
      c = uint8_t((((const uint32_t *)pixels)[i] / maxVal) * 255);
    } else if (componentType == DVR_IMAGE_SRGB8)
      c = uint8_t(glm::convertSRGBToLinear(glm::vec1(((const uint8_t *)pixels)[i] / 255.f)).x * 255);
    else
      c = ((const uint8_t *)pixels)[i];
    staging[o++] = c;
    if (channels == 3 && o % 4 == 3)
      staging[o++] = 255;
  }
  if (nc == 3)
    nc = 4;
  auto *img = new RefImage();
  auto desc = cudaCreateChannelDesc(nc >= 1 ? 8 : 0, nc >= 2 ? 8 : 0, nc >= 3 ? 8 : 0, nc >= 4 ? 8 : 0,
      cudaChannelFormatKindUnsigned);
  RCK(cudaMalloc3DArray(&img->arr, &desc, make_cudaExtent(w, h, 0)));
  cudaMemcpy3DParms p = {};
  p.dstArray = img->arr;
  p.srcPtr = make_cudaPitchedPtr(staging.data(), w * nc * sizeof(uint8_t), w, h);
  p.extent = make_cudaExtent(w, h, 1);
  p.kind = cudaMemcpyHostToDevice;
  RCK(cudaMemcpy3D(&p));
  cudaResourceDesc rd;
  std::memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = img->arr;
  cudaTextureDesc td;
  std::memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = true;
  RCK(cudaCreateTextureObject(&img->tex, &rd, &td, nullptr));
  *out = img;
  return 0;
}

int refgpu_image_destroy(RefImage *img)
{
  if (!img) return 0;
  if (img->tex) cudaDestroyTextureObject(img->tex);
  if (img->arr) cudaFreeArray(img->arr);
  delete img;
  return 0;
}

// Renderer::populateFrameData (Renderer.cpp:191-198): an image switches the frame to BackgroundMode::IMAGE
int refgpu_scene_set_background_image(RefScene *s, RefImage *img)
{
  s->bgTex = img ? img->tex : 0;
  return 0;
}

int refgpu_scene_create(const RefInstance *inst, int n, RefScene **out)
{
  auto *s = new RefScene();
  s->n = n;
  std::vector<SpatialFieldGPUData> fields(n);
  std::vector<VolumeGPUData> vols(n);
  std::vector<InstanceVolumeGPUData> insts(n);
  std::vector<DeviceObjectIndex> idx(n);
  std::vector<RefInstanceXfm> xf(n);
  RCK(cudaMalloc(&s->volIdx, sizeof(DeviceObjectIndex) * (n ? n : 1)));
  static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int i = 0; i < n; ++i) {
    fields[i] = inst[i].volume->field->gpu;
    fields[i].grid.dims = inst[i].volume->gridDims;
    fields[i].grid.worldBounds = inst[i].volume->field->bounds;
    fields[i].grid.maxOpacities = inst[i].volume->maxOpacities;
    if (!inst[i].volume->maxOpacities)
      s->hasGrid = false;
    vols[i] = inst[i].volume->gpu;
    vols[i].data.tf1d.field = i;
    idx[i] = i;
    insts[i].volumes = s->volIdx + i;
    insts[i].id = inst[i].instanceId;
    std::memcpy(xf[i].m, inst[i].worldToObject, sizeof(ident));
    xf[i].identity = std::memcmp(xf[i].m, ident, sizeof(ident)) == 0;
  }
  const int m = n ? n : 1;
  RCK(cudaMalloc(&s->fields, sizeof(SpatialFieldGPUData) * m));
  RCK(cudaMalloc(&s->volumes, sizeof(VolumeGPUData) * m));
  RCK(cudaMalloc(&s->instances, sizeof(InstanceVolumeGPUData) * m));
  RCK(cudaMalloc(&s->xfms, sizeof(RefInstanceXfm) * m));
  RCK(cudaMalloc(&s->camera, sizeof(CameraGPUData)));
  if (n) {
    RCK(cudaMemcpy(s->fields, fields.data(), sizeof(SpatialFieldGPUData) * n, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(s->volumes, vols.data(), sizeof(VolumeGPUData) * n, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(s->instances, insts.data(), sizeof(InstanceVolumeGPUData) * n, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(s->volIdx, idx.data(), sizeof(DeviceObjectIndex) * n, cudaMemcpyHostToDevice));
    RCK(cudaMemcpy(s->xfms, xf.data(), sizeof(RefInstanceXfm) * n, cudaMemcpyHostToDevice));
  }
  *out = s;
  return 0;
}

int refgpu_scene_destroy(RefScene *s)
{
  if (!s) return 0;
  cudaFree(s->fields); cudaFree(s->volumes); cudaFree(s->instances); cudaFree(s->volIdx); cudaFree(s->xfms);
  cudaFree(s->camera);
  delete s;
  return 0;
}

// Frame::renderFrame: newFrame() resets + upload() + launch (Frame.cu:272-289)
int refgpu_render(const DvrFrameParams *p, const DvrCamera *c, RefScene *scene, const DvrFrameBuffers *b,
    void *stream)
{
  cudaStream_t s = (cudaStream_t)stream;
  CameraGPUData cam{};
  cam.type = c->type == DVR_CAMERA_PERSPECTIVE ? CameraType::PERSPECTIVE : CameraType::ORTHOGRAPHIC;
  cam.region = vec4(c->region[0], c->region[1], c->region[2], c->region[3]);
  cam.pos = vec3(c->pos[0], c->pos[1], c->pos[2]);
  cam.dir = vec3(c->dir[0], c->dir[1], c->dir[2]);
  cam.up = vec3(c->up[0], c->up[1], c->up[2]);
  if (c->type == DVR_CAMERA_PERSPECTIVE) {
    cam.perspective.dir_du = vec3(c->du[0], c->du[1], c->du[2]);
    cam.perspective.dir_dv = vec3(c->dv[0], c->dv[1], c->dv[2]);
    cam.perspective.dir_00 = vec3(c->p00[0], c->p00[1], c->p00[2]);
    cam.perspective.scaledAperture = c->scaledAperture;
    cam.perspective.aspect = c->aspect;
  } else {
    cam.orthographic.pos_du = vec3(c->du[0], c->du[1], c->du[2]);
    cam.orthographic.pos_dv = vec3(c->dv[0], c->dv[1], c->dv[2]);
    cam.orthographic.pos_00 = vec3(c->p00[0], c->p00[1], c->p00[2]);
  }
  RCK(cudaMemcpyAsync(scene->camera, &cam, sizeof(cam), cudaMemcpyHostToDevice, s));

  FrameGPUData fd;
  std::memset(&fd, 0, sizeof(fd));
  fd.fb.buffers.colorAccumulation = (vec4 *)b->colorAccumulation;
  if (p->format == DVR_FORMAT_FLOAT32_VEC4)
    fd.fb.buffers.outColorVec4 = (vec4 *)b->outColor;
  else
    fd.fb.buffers.outColorUint = (uint32_t *)b->outColor;
  fd.fb.buffers.depth = b->depth;
  fd.fb.buffers.primID = b->primId;
  fd.fb.buffers.objID = b->objId;
  fd.fb.buffers.instID = b->instId;
  fd.fb.buffers.albedo = (vec3 *)b->albedo;
  fd.fb.buffers.normal = (vec3 *)b->normal;
  fd.fb.frameID = p->frameID;
  fd.fb.checkerboardID = p->checkerboardID;
  fd.fb.invFrameID = 1.f / (p->frameID + 1);
  fd.fb.format = p->format == DVR_FORMAT_FLOAT32_VEC4
      ? FrameFormat::FLOAT
      : (p->format == DVR_FORMAT_UFIXED8_RGBA_SRGB ? FrameFormat::SRGB : FrameFormat::UINT);
  fd.fb.size = uvec2(p->width, p->height);
  fd.fb.invSize = 1.f / vec2(fd.fb.size);
  if (scene->bgTex) {
    fd.renderer.backgroundMode = BackgroundMode::IMAGE;
    fd.renderer.background.texobj = scene->bgTex;
  } else {
    fd.renderer.backgroundMode = BackgroundMode::COLOR;
    fd.renderer.background.color = vec4(p->background[0], p->background[1], p->background[2], p->background[3]);
  }
  fd.renderer.ambientColor = vec3(1.f);
  fd.renderer.ambientIntensity = p->ambientRadiance;
  fd.renderer.occlusionDistance = p->occlusionDistance > 0.f ? p->occlusionDistance : 1e20f;
  fd.renderer.params.dpt.maxDepth = p->maxDepth <= 0 ? 5 : (p->maxDepth > 256 ? 256 : p->maxDepth);
  fd.renderer.cullTriangleBF = false;
  fd.renderer.inverseVolumeSamplingRate = p->inverseVolumeSamplingRate;
  fd.renderer.numIterations = p->checkerboardID >= 0 ? 1 : (p->numIterations > 1 ? p->numIterations : 1);
  fd.renderer.maxRayDepth = 5;
  fd.world.volumeInstances = scene->instances;
  fd.world.numVolumeInstances = scene->n;
  fd.world.hdri = -1;
  fd.camera = scene->camera;
  fd.registry.fields = scene->fields;
  fd.registry.volumes = scene->volumes;

  const size_t npx = (size_t)p->width * p->height;
  if (p->frameID == 0 && p->checkerboardID <= 0) { // Frame::newFrame reset, Frame.cu:594-647
    fill((vec4 *)b->colorAccumulation, npx, vec4(0.f), s);
    fill(b->depth, npx, std::numeric_limits<float>::max(), s);
    fill(b->primId, npx, uint32_t(0), s);
    fill(b->objId, npx, uint32_t(0), s);
    fill(b->instId, npx, uint32_t(0), s);
    fill((vec3 *)b->albedo, npx, vec3(0.f), s);
    fill((vec3 *)b->normal, npx, vec3(0.f), s);
  }
  RCK(cudaMemcpyToSymbolAsync(frameData, &fd, sizeof(fd), 0, cudaMemcpyHostToDevice, s));
  const RefInstanceXfm *xf = scene->xfms;
  RCK(cudaMemcpyToSymbolAsync(g_xfms, &xf, sizeof(xf), 0, cudaMemcpyHostToDevice, s));
  const uint32_t lw = p->checkerboardID >= 0 ? (p->width + 1) / 2 : p->width;
  const uint32_t lh = p->checkerboardID >= 0 ? (p->height + 1) / 2 : p->height;
  dim3 block(16, 8), grid((lw + 15) / 16, (lh + 7) / 8);
  if (p->integrator == DVR_INTEGRATOR_TEST)
    refgpu_raygen_test<<<grid, block, 0, s>>>();
  else if (p->integrator == DVR_INTEGRATOR_DPT) {
    for (int i = 0; i < scene->n; ++i)
      if (!scene->hasGrid) {
        snprintf(g_err, sizeof(g_err), "dpt needs refgpu_volume_set_grid on every volume");
        return -1;
      }
    refgpu_raygen_dpt<<<grid, block, 0, s>>>();
  } else
    refgpu_raygen<<<grid, block, 0, s>>>(p->integrator == DVR_INTEGRATOR_RAYCAST ? 1 : 0);
  RCK(cudaGetLastError());
  return 0;
}

} // extern "C"
