// device.cpp — Frame, Device and the extern "C" ANARI entry points of libanari_library_visrtx_b200.so.
//
// Frame follows frame/Frame.cu:75-193,195-310,312-423,574-660 of the reference (channel allocation,
// accumulation-reset bookkeeping, asynchronous render on the device's private stream, event timing,
// host / CUDA mapping); Device follows VisRTXDevice.cpp:386-473,554-640 (lazy init, `cudaDevice`,
// `forceInit`, sticky failure, status callback).  The only GPU work issued here besides memory
// management is dvr_render() / dvr_scale_vec3() from include/dvr_b200.h.
#include "objects.h"

#include <anari/ext/visrtx_b200.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace b200 {

// ---------------------------------------------------------------------------------------------------------
// CudaDeviceScope
// ---------------------------------------------------------------------------------------------------------
CudaDeviceScope::CudaDeviceScope(Device *d)
{
  if (!d || d->cudaDevice() < 0)
    return;
  if (cudaGetDevice(&prev) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  if (prev != d->cudaDevice()) {
    cudaSetDevice(d->cudaDevice());
    active = true;
  }
}
CudaDeviceScope::~CudaDeviceScope()
{
  if (active)
    cudaSetDevice(prev);
}

GpuScope::GpuScope(int gpu)
{
  if (cudaGetDevice(&prev) != cudaSuccess) {
    cudaGetLastError();
    prev = -1;
  }
  if (prev != gpu)
    cudaSetDevice(gpu);
  else
    prev = -1;
}
GpuScope::~GpuScope()
{
  if (prev >= 0)
    cudaSetDevice(prev);
}

// ---------------------------------------------------------------------------------------------------------
// Frame
// ---------------------------------------------------------------------------------------------------------
Frame::Frame(Device *d) : Object(d, ANARI_FRAME) {}

Frame::~Frame()
{
  if (m_eventEnd)
    wait();
  syncAllGpus();
  freeMultiGpuBuffers();
  CudaDeviceScope scope(device);
  freeBuffers();
  if (m_eventStart) cudaEventDestroy((cudaEvent_t)m_eventStart);
  if (m_eventEnd) cudaEventDestroy((cudaEvent_t)m_eventEnd);
  if (m_pinned) cudaFreeHost(m_pinned);
}

void Frame::freeBuffers()
{
  void **bufs[] = {&m_accum, &m_color, &m_depth, &m_primId, &m_objId, &m_instId, &m_albedoAccum, &m_normalAccum,
      &m_albedo, &m_normal};
  for (void **b : bufs) {
    if (*b)
      cudaFree(*b);
    *b = nullptr;
  }
}

// ---- multi-GPU frames -------------------------------------------------------------------------------------
// One process drives every GPU of the device (peer access enabled in Device::initDevice).  The display GPU (rank 0)
// owns the output channels the application maps; every GPU owns the accumulation state of the pixels it resolves.
namespace {
constexpr uint32_t kMaxRegions = 1024;                       // rows of a region-flag table (DvrSlabExchange)
constexpr size_t kFlagWords = (size_t)kMaxRegions * 16 + 32; // region table, resolved table, error word
} // namespace

void Frame::syncAllGpus() const
{
  if (!device || device->gpuCount() <= 1)
    return;
  for (int r = 0; r < device->gpuCount(); ++r) {
    GpuScope scope(device->gpu(r));
    cudaStreamSynchronize((cudaStream_t)device->stream(r));
  }
}

void Frame::freeMultiGpuBuffers()
{
  for (size_t r = 0; r < m_perGpu.size(); ++r) {
    GpuScope scope(device->gpu((int)r));
    PerGpu &g = m_perGpu[r];
    if (r > 0) { // rank 0 uses the frame's own accumulation / depth buffers
      cudaFree(g.accum);
      cudaFree(g.depth);
    }
    cudaFree(g.partial[0]);
    cudaFree(g.partial[1]);
    cudaFree(g.flags);
    cudaFree(g.regionDone);
    if (g.done)
      cudaEventDestroy((cudaEvent_t)g.done);
  }
  m_perGpu.clear();
  m_seq = 0;
}

bool Frame::ensureMultiGpuBuffers()
{
  const int world = device->gpuCount();
  if (world <= 1)
    return false;
  if ((int)m_perGpu.size() == world)
    return true;
  const size_t n = (size_t)m_size[0] * m_size[1];
  m_perGpu.assign((size_t)world, PerGpu());
  bool ok = true;
  for (int r = 0; r < world && ok; ++r) {
    GpuScope scope(device->gpu(r));
    PerGpu &g = m_perGpu[(size_t)r];
    if (r == 0) {
      g.accum = m_accum;
      g.depth = m_depth;
    } else {
      ok = ok && cudaMalloc(&g.accum, n * 16) == cudaSuccess;
      if (m_depth)
        ok = ok && cudaMalloc(&g.depth, n * 4) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&g.partial[0], n * 20) == cudaSuccess && cudaMalloc(&g.partial[1], n * 20) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&g.flags, kFlagWords * 4) == cudaSuccess
        && cudaMalloc((void **)&g.regionDone, kMaxRegions * 4) == cudaSuccess;
    ok = ok && cudaMemset(g.flags, 0, kFlagWords * 4) == cudaSuccess
        && cudaMemset(g.regionDone, 0, kMaxRegions * 4) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags((cudaEvent_t *)&g.done, cudaEventDisableTiming) == cudaSuccess;
    cudaDeviceSynchronize();
  }
  if (!ok) {
    cudaGetLastError();
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_OUT_OF_MEMORY,
        "multi-GPU frame buffers could not be allocated; rendering on the display GPU only");
    freeMultiGpuBuffers();
    return false;
  }
  m_seq = 0;
  return true;
}

// Sort-last: the volume's z-slabs live on the GPUs; ONE fused launch per GPU marches its slab, exchanges region flags
// and composites + resolves the regions it owns straight into the display GPU's channels (dvr_render_slab_frame).
bool Frame::renderSortLast(const DvrFrameParams &p, const DvrFrameBuffers &display, Volume *v, const FlatInstance &fi)
{
  const int world = device->gpuCount();
  const size_t n = (size_t)m_size[0] * m_size[1];
  const uint32_t seq = ++m_seq;
  const int parity = (int)(seq & 1u);
  const float *rgba[16];
  const float *depth[16];
  unsigned int *regionFlags[16], *resolvedFlags[16];
  for (int q = 0; q < world; ++q) {
    rgba[q] = (const float *)m_perGpu[(size_t)q].partial[parity];
    depth[q] = (const float *)((const uint8_t *)m_perGpu[(size_t)q].partial[parity] + n * 16);
    regionFlags[q] = m_perGpu[(size_t)q].flags;
    resolvedFlags[q] = m_perGpu[(size_t)q].flags + (size_t)kMaxRegions * 16;
  }
  bool ok = true;
  for (int k = 0; k < world; ++k) {
    const int r = (k + 1) % world; // the display GPU's launch waits for all the others: enqueue it last
    GpuScope scope(device->gpu(r));
    DvrFrameBuffers b = display; // colour / ids: peer stores into the display GPU's channels
    b.colorAccumulation = (float *)m_perGpu[(size_t)r].accum;
    b.depth = (float *)m_perGpu[(size_t)r].depth;
    b.depthMirror = (r != 0 && display.depth) ? display.depth : nullptr;
    DvrVolumeInstance inst;
    inst.volume = v->part(r);
    std::memcpy(inst.worldToObject, fi.worldToObject, sizeof(inst.worldToObject));
    inst.instanceId = fi.instId;
    inst._pad = 0;
    DvrSlabExchange x;
    std::memset(&x, 0, sizeof(x));
    x.nRanks = (uint32_t)world;
    x.rank = (uint32_t)r;
    x.seq = seq;
    x.maxRegions = kMaxRegions;
    x.partialRgba = rgba;
    x.partialDepth = depth;
    x.regionFlags = regionFlags;
    x.resolvedFlags = resolvedFlags;
    x.regionDone = m_perGpu[(size_t)r].regionDone;
    x.errorFlag = m_perGpu[(size_t)r].flags + (size_t)kMaxRegions * 16 + 16;
    x.waitAllResolved = r == 0 ? 1 : 0;
    const int rc = dvr_render_slab_frame(&p, &m_camera->cam, &inst, v->id(), fi.instId, &b, &x, device->stream(r));
    if (rc != DVR_OK) {
      report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "dvr_render_slab_frame failed on GPU %d: %s",
          device->gpu(r), dvr_last_error());
      ok = false;
    }
  }
  return ok;
}

// Sort-first: every GPU holds every field; GPU r renders the tile rows (row % N == r) and stores the colour / ids of
// its pixels into the display GPU's channels; the display stream then waits for the others.
bool Frame::renderSortFirst(DvrFrameParams p, const DvrFrameBuffers &display, const std::vector<FlatInstance> &flat)
{
  const int world = device->gpuCount();
  bool ok = true;
  for (int k = 0; k < world; ++k) {
    const int r = (k + 1) % world;
    GpuScope scope(device->gpu(r));
    std::vector<DvrVolumeInstance> inst(flat.size());
    for (size_t i = 0; i < flat.size(); ++i) {
      inst[i].volume = flat[i].volume->part(r);
      std::memcpy(inst[i].worldToObject, flat[i].worldToObject, sizeof(inst[i].worldToObject));
      inst[i].instanceId = flat[i].instId;
      inst[i]._pad = 0;
    }
    DvrFrameBuffers b = display;
    b.colorAccumulation = (float *)m_perGpu[(size_t)r].accum;
    b.depth = (float *)m_perGpu[(size_t)r].depth;
    b.depthMirror = (r != 0 && display.depth) ? display.depth : nullptr;
    p.tileRank = (uint32_t)r;
    p.tileRanks = (uint32_t)world;
    cudaStream_t s = (cudaStream_t)device->stream(r);
    const int rc = dvr_render(&p, &m_camera->cam, inst.data(), (uint32_t)inst.size(), &b, s);
    if (rc != DVR_OK) {
      report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "dvr_render failed on GPU %d: %s", device->gpu(r),
          dvr_last_error());
      ok = false;
    }
    if (r != 0) {
      cudaEventRecord((cudaEvent_t)m_perGpu[(size_t)r].done, s);
      GpuScope display0(device->gpu(0));
      cudaStreamWaitEvent((cudaStream_t)device->stream(0), (cudaEvent_t)m_perGpu[(size_t)r].done, 0);
    }
  }
  return ok;
}

bool Frame::isValid() const
{
  return m_valid && m_renderer && m_renderer->isValid() && m_camera && m_camera->isValid() && m_world
      && m_world->isValid();
}

void Frame::commitParameters()
{
  m_renderer.reset(static_cast<Renderer *>(getParamObject("renderer", ANARI_RENDERER)));
  m_camera.reset(static_cast<Camera *>(getParamObject("camera", ANARI_CAMERA)));
  m_world.reset(static_cast<World *>(getParamObject("world", ANARI_WORLD)));
  m_callback = getParam<ANARIFrameCompletionCallback>("frameCompletionCallback", ANARI_FRAME_COMPLETION_CALLBACK, nullptr);
  m_callbackUserPtr = getParam<const void *>("frameCompletionCallbackUserData", ANARI_VOID_POINTER, nullptr);
  m_colorType = getParam<ANARIDataType>("channel.color", ANARI_DATA_TYPE, ANARI_UFIXED8_RGBA_SRGB);
  m_size[0] = m_size[1] = 10;
  getParamRaw("size", ANARI_UINT32_VEC2, m_size, sizeof(m_size));
  m_depthType = getParam<ANARIDataType>("channel.depth", ANARI_DATA_TYPE, ANARI_UNKNOWN);
  m_primIdType = getParam<ANARIDataType>("channel.primitiveId", ANARI_DATA_TYPE, ANARI_UNKNOWN);
  m_objIdType = getParam<ANARIDataType>("channel.objectId", ANARI_DATA_TYPE, ANARI_UNKNOWN);
  m_instIdType = getParam<ANARIDataType>("channel.instanceId", ANARI_DATA_TYPE, ANARI_UNKNOWN);
  m_albedoType = getParam<ANARIDataType>("channel.albedo", ANARI_DATA_TYPE, ANARI_UNKNOWN);
  m_normalType = getParam<ANARIDataType>("channel.normal", ANARI_DATA_TYPE, ANARI_UNKNOWN);
}

void Frame::finalize()
{
  m_valid = false;
  if (!(m_renderer && m_camera && m_world))
    return;
  if (!device->initDevice())
    return;
  if (m_eventEnd)
    wait();
  syncAllGpus();
  freeMultiGpuBuffers();
  CudaDeviceScope scope(device);
  freeBuffers();
  m_pinnedHoldsFrame = false;

  if (m_colorType == ANARI_FLOAT32_VEC4)
    m_format = DVR_FORMAT_FLOAT32_VEC4;
  else if (m_colorType == ANARI_UFIXED8_RGBA_SRGB)
    m_format = DVR_FORMAT_UFIXED8_RGBA_SRGB;
  else
    m_format = DVR_FORMAT_UFIXED8_VEC4;

  const bool chPrim = m_primIdType == ANARI_UINT32, chObj = m_objIdType == ANARI_UINT32,
             chInst = m_instIdType == ANARI_UINT32;
  const bool chAlbedo = m_albedoType == ANARI_FLOAT32_VEC3 || m_albedoType == ANARI_FLOAT32;
  const bool chNormal = m_normalType == ANARI_FLOAT32_VEC3 || m_normalType == ANARI_FLOAT32;
  const bool chDepth = m_depthType == ANARI_FLOAT32 || chPrim || chObj || chInst; // Frame.cu:120-123
  if (chDepth && m_depthType != ANARI_FLOAT32)
    m_depthType = ANARI_FLOAT32;

  const size_t n = (size_t)m_size[0] * m_size[1];
  auto alloc = [&](void **p, size_t bytes) -> bool {
    if (bytes == 0)
      return true;
    if (cudaMalloc(p, bytes) != cudaSuccess) {
      cudaGetLastError();
      report(ANARI_SEVERITY_ERROR, ANARI_STATUS_OUT_OF_MEMORY, "frame buffer allocation of %zu bytes failed", bytes);
      return false;
    }
    return true;
  };
  bool ok = n > 0;
  ok = ok && alloc(&m_accum, n * 16);
  ok = ok && alloc(&m_color, n * (m_format == DVR_FORMAT_FLOAT32_VEC4 ? 16 : 4));
  ok = ok && alloc(&m_depth, chDepth ? n * 4 : 0);
  ok = ok && alloc(&m_primId, chPrim ? n * 4 : 0);
  ok = ok && alloc(&m_objId, chObj ? n * 4 : 0);
  ok = ok && alloc(&m_instId, chInst ? n * 4 : 0);
  ok = ok && alloc(&m_albedoAccum, chAlbedo ? n * 12 : 0);
  ok = ok && alloc(&m_albedo, chAlbedo ? n * 12 : 0);
  ok = ok && alloc(&m_normalAccum, chNormal ? n * 12 : 0);
  ok = ok && alloc(&m_normal, chNormal ? n * 12 : 0);
  if (!ok) {
    freeBuffers();
    return;
  }
  if (!m_eventStart) {
    cudaEventCreate((cudaEvent_t *)&m_eventStart);
    cudaEventCreate((cudaEvent_t *)&m_eventEnd);
    cudaEventRecord((cudaEvent_t)m_eventStart, (cudaStream_t)device->stream());
    cudaEventRecord((cudaEvent_t)m_eventEnd, (cudaStream_t)device->stream());
  }
  m_valid = true;
  m_nextFrameReset = true;
}

void Frame::checkAccumulationReset()
{ // Frame.cu:574-588: anything finalised since the last frame restarts accumulation
  if (m_nextFrameReset)
    return;
  if (m_lastCommitSeen < device->lastFinalization()) {
    m_lastCommitSeen = device->lastFinalization();
    m_nextFrameReset = true;
  }
}

void Frame::wait() const
{
  if (m_eventEnd)
    cudaEventSynchronize((cudaEvent_t)m_eventEnd);
}

int Frame::ready(ANARIWaitMask m)
{
  if (!m_eventEnd)
    return 1;
  if (m == ANARI_NO_WAIT)
    return cudaEventQuery((cudaEvent_t)m_eventEnd) == cudaSuccess;
  wait();
  return 1;
}

void Frame::renderFrame()
{
  wait();
  device->flushCommits();
  if (!isValid()) {
    const char *problem = "<unknown>";
    if (!m_renderer) problem = "missing ANARIRenderer";
    else if (!m_renderer->isValid()) problem = "invalid ANARIRenderer";
    else if (!m_camera) problem = "missing ANARICamera";
    else if (!m_camera->isValid()) problem = "invalid ANARICamera";
    else if (!m_world) problem = "missing ANARIWorld";
    else if (!m_world->isValid()) problem = "invalid ANARIWorld";
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_OPERATION,
        "skipping render of incomplete or invalid frame object -- issue: %s", problem);
    return;
  }
  CudaDeviceScope scope(device);
  cudaStream_t stream = (cudaStream_t)device->stream();

  if (m_lastCommitSeen == 0)
    m_lastCommitSeen = device->lastFinalization();
  checkAccumulationReset();
  m_lastCommitSeen = device->lastFinalization();

  const int sampleLimit = m_renderer->sampleLimit;
  if (!m_nextFrameReset && sampleLimit > 0 && m_frameID >= sampleLimit)
    return; // Frame.cu:251-253

  cudaEventRecord((cudaEvent_t)m_eventStart, stream);

  // Frame::newFrame, Frame.cu:590-660 (buffer clears are folded into the launch: frameID == 0)
  const bool cb = m_renderer->checkerboard;
  if (m_nextFrameReset) {
    m_frameID = 0;
    m_checkerboardID = cb ? 0 : -1;
    m_nextFrameReset = false;
  } else {
    if (cb)
      m_frameID += int(m_checkerboardID == 3);
    else
      m_frameID += m_renderer->spp;
    m_checkerboardID = cb ? ((m_checkerboardID + 1) & 0x3) : -1;
  }
  m_invFrameID = 1.f / (m_frameID + 1);

  const std::vector<FlatInstance> flat = m_world->flatten(true);
  // surfaces in the world: the frame takes the mixed-scene launch (marching renderers only)
  DvrSurfaces *surfaceSet = nullptr;
  if (m_renderer->integrator == DVR_INTEGRATOR_RAYCAST || m_renderer->integrator == DVR_INTEGRATOR_DEFAULT)
    surfaceSet = m_world->surfaceSet(true);
  else if (m_world->surfaceSet(false))
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_OPERATION,
        "the '%s' renderer of this device draws volumes only: the world's surfaces are ignored", m_renderer->subtype.c_str());

  DvrFrameParams p;
  std::memset(&p, 0, sizeof(p));
  p.width = m_size[0];
  p.height = m_size[1];
  p.format = m_format;
  p.integrator = m_renderer->integrator;
  p.frameID = m_frameID;
  p.checkerboardID = m_checkerboardID;
  p.numIterations = std::max(m_renderer->spp, 1);
  p.inverseVolumeSamplingRate = 1.f / m_renderer->volumeSamplingRate;
  std::memcpy(p.background, m_renderer->background, sizeof(p.background));
  p.backgroundImage = m_renderer->backgroundImage();
  p.tileRank = m_renderer->tileRank;
  p.tileRanks = m_renderer->tileRanks;
  p.useMacrocellSkipping = m_renderer->macrocellSkipping;
  p.maxDepth = m_renderer->maxDepth;
  p.dptReferenceGrid = m_renderer->dptReferenceGrid ? 1 : 0;
  p.ambientRadiance = m_renderer->ambientRadiance;
  p.occlusionDistance = m_renderer->occlusionDistance;

  DvrFrameBuffers b;
  std::memset(&b, 0, sizeof(b));
  b.colorAccumulation = (float *)m_accum;
  b.outColor = m_color;
  b.outColorMirror = nullptr;
  m_pinnedHoldsFrame = false;
  if (m_streamColorToHost
      && ensurePinned((size_t)m_size[0] * m_size[1] * (m_format == DVR_FORMAT_FLOAT32_VEC4 ? 16 : 4))) {
    b.outColorMirror = m_pinned;
    m_pinnedHoldsFrame = true;
  }
  b.depth = (float *)m_depth;
  b.primId = (uint32_t *)m_primId;
  b.objId = (uint32_t *)m_objId;
  b.instId = (uint32_t *)m_instId;
  b.albedo = (float *)m_albedoAccum;
  b.normal = (float *)m_normalAccum;

  // Multi-GPU device: the frame goes to all GPUs when the scene is one the distributed paths cover — marching
  // renderer, one sample per pixel-pass, no per-GPU auxiliary accumulations (albedo / normal) and no background
  // texture (texture objects are per GPU); sort-last additionally needs a single slabbed structuredRegular volume.
  bool done = false;
  if (surfaceSet) { // mixed scene: one GPU (the display GPU), dvr_render_scene
    std::vector<DvrVolumeInstance> inst;
    for (size_t i = 0; i < flat.size(); ++i) {
      DvrVolumeInstance in;
      in.volume = flat[i].volume->whole();
      if (!in.volume)
        continue;
      std::memcpy(in.worldToObject, flat[i].worldToObject, sizeof(in.worldToObject));
      in.instanceId = flat[i].instId;
      in._pad = 0;
      inst.push_back(in);
    }
    const std::vector<DvrLight> lights = m_world->flattenLights();
    DvrSceneParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.surfaces = surfaceSet;
    sp.lights = lights.data();
    sp.nLights = (uint32_t)lights.size();
    std::memcpy(sp.ambientColor, m_renderer->ambientColor, sizeof(sp.ambientColor));
    sp.ambientRadiance = m_renderer->ambientRadiance;
    sp.occlusionDistance = m_renderer->occlusionDistance;
    sp.ambientSamples = m_renderer->ambientSamples;
    sp.cullTriangleBackfaces = m_renderer->cullTriangleBackfaces ? 1 : 0;
    const int rc = dvr_render_scene(&p, &m_camera->cam, inst.data(), (uint32_t)inst.size(), &sp, &b, stream);
    if (rc != DVR_OK)
      report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "dvr_render_scene failed: %s", dvr_last_error());
    done = true;
  } else if (device->gpuCount() > 1) {
    const bool marching = p.integrator == DVR_INTEGRATOR_RAYCAST || p.integrator == DVR_INTEGRATOR_DEFAULT;
    const bool plain = marching && !m_albedoAccum && !m_normalAccum && !p.backgroundImage && p.tileRanks <= 1;
    bool allDistributed = !flat.empty();
    for (const FlatInstance &fi : flat)
      allDistributed = allDistributed && fi.volume->distributed();
    if (plain && allDistributed && device->sortLast() && flat.size() == 1 && flat[0].volume->field()->slabbed()
        && p.numIterations == 1 && p.checkerboardID < 0 && ensureMultiGpuBuffers())
      done = renderSortLast(p, b, flat[0].volume, flat[0]);
    else if (plain && allDistributed && device->sortFirst() && ensureMultiGpuBuffers())
      done = renderSortFirst(p, b, flat);
  }
  if (!done) {
    std::vector<DvrVolumeInstance> inst;
    inst.reserve(flat.size());
    for (size_t i = 0; i < flat.size(); ++i) {
      DvrVolumeInstance in;
      in.volume = flat[i].volume->whole(); // (multi-GPU device: uploads the whole field to the display GPU on demand)
      if (!in.volume)
        continue;
      std::memcpy(in.worldToObject, flat[i].worldToObject, sizeof(in.worldToObject));
      in.instanceId = flat[i].instId;
      in._pad = 0;
      inst.push_back(in);
    }
    const int rc = dvr_render(&p, &m_camera->cam, inst.data(), (uint32_t)inst.size(), &b, stream);
    if (rc != DVR_OK)
      report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "dvr_render failed: %s", dvr_last_error());
  }
  m_everRendered = true;

  if (m_callback) { // Frame.cu:295-304
    cudaLaunchHostFunc(stream,
        [](void *self_) {
          Frame *self = (Frame *)self_;
          self->m_callback(self->m_callbackUserPtr, (ANARIDevice)self->device, (ANARIFrame)self);
        },
        this);
  }
  cudaEventRecord((cudaEvent_t)m_eventEnd, stream);
}

bool Frame::ensurePinned(size_t bytes)
{
  if (m_pinnedBytes >= bytes && m_pinned)
    return true;
  if (m_pinned)
    cudaFreeHost(m_pinned);
  m_pinned = nullptr;
  m_pinnedBytes = 0;
  // portable + mapped: every GPU of a multi-GPU device streams its pixels into it
  if (cudaHostAlloc(&m_pinned, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    m_pinned = nullptr;
    return false;
  }
  m_pinnedBytes = bytes;
  return true;
}

void *Frame::download(void *dev, size_t bytes, std::vector<uint8_t> &host)
{
  if (!dev)
    return nullptr;
  CudaDeviceScope scope(device);
  if (&host == &m_hColor) { // colour is mapped every frame: hand out the pinned buffer directly
    if (m_pinnedHoldsFrame && m_pinned)
      return m_pinned; // the launch already streamed this frame's colour to the host (wait() has run)
    m_streamColorToHost = true; // from the next launch on
    if (ensurePinned(bytes)) {
      cudaMemcpyAsync(m_pinned, dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)device->stream());
      cudaStreamSynchronize((cudaStream_t)device->stream());
      m_pinnedHoldsFrame = true; // until the next launch
      return m_pinned;
    }
  }
  host.resize(bytes);
  cudaMemcpy(host.data(), dev, bytes, cudaMemcpyDeviceToHost);
  return host.data();
}

const void *Frame::map(const std::string &channelIn, uint32_t *w, uint32_t *h, ANARIDataType *pixelType)
{ // Frame.cu:312-423
  wait();
  std::string channel = channelIn;
  bool gpu = false;
  auto endsWith = [&](const char *suf) {
    const size_t n = std::strlen(suf);
    return channel.size() > n && channel.compare(channel.size() - n, n, suf) == 0;
  };
  if (endsWith("CUDA")) {
    gpu = true;
    channel.resize(channel.size() - 4);
  } else if (endsWith("GPU")) {
    gpu = true;
    channel.resize(channel.size() - 3);
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_NO_ERROR, "%s is deprecated, please use %sCUDA instead",
        channelIn.c_str(), channel.c_str());
  }
  const size_t n = (size_t)m_size[0] * m_size[1];
  ANARIDataType type = ANARI_UNKNOWN;
  void *ret = nullptr;
  if (!m_valid) {
    // fall through: unknown
  } else if (channel == "channel.color") {
    type = m_colorType;
    ret = gpu ? m_color : download(m_color, n * (m_format == DVR_FORMAT_FLOAT32_VEC4 ? 16 : 4), m_hColor);
  } else if (channel == "channel.depth" && m_depthType == ANARI_FLOAT32 && m_depth) {
    type = ANARI_FLOAT32;
    ret = gpu ? m_depth : download(m_depth, n * 4, m_hDepth);
  } else if (channel == "channel.primitiveId" && m_primId) {
    type = ANARI_UINT32;
    ret = gpu ? m_primId : download(m_primId, n * 4, m_hPrim);
  } else if (channel == "channel.objectId" && m_objId) {
    type = ANARI_UINT32;
    ret = gpu ? m_objId : download(m_objId, n * 4, m_hObj);
  } else if (channel == "channel.instanceId" && m_instId) {
    type = ANARI_UINT32;
    ret = gpu ? m_instId : download(m_instId, n * 4, m_hInst);
  } else if ((channel == "channel.albedo" && m_albedo) || (channel == "channel.normal" && m_normal)) {
    // Frame::mapAlbedoBuffer / mapNormalBuffer: averaged at map time (Frame.cu:521-557)
    const bool alb = channel == "channel.albedo";
    CudaDeviceScope scope(device);
    dvr_scale_vec3((const float *)(alb ? m_albedoAccum : m_normalAccum), (float *)(alb ? m_albedo : m_normal), n,
        m_invFrameID, device->stream());
    cudaStreamSynchronize((cudaStream_t)device->stream());
    type = ANARI_FLOAT32_VEC3;
    void *dev = alb ? m_albedo : m_normal;
    ret = gpu ? dev : download(dev, n * 12, alb ? m_hAlbedo : m_hNormal);
  }
  if (type != ANARI_UNKNOWN) {
    if (w) *w = m_size[0];
    if (h) *h = m_size[1];
  }
  if (pixelType)
    *pixelType = type;
  return ret;
}

bool Frame::getProperty(const std::string &name, ANARIDataType t, void *mem, uint64_t size, uint32_t mask)
{ // Frame.cu:163-193
  if (t == ANARI_FLOAT32 && name == "duration" && size >= 4) {
    if (mask & ANARI_WAIT)
      wait();
    float ms = 0.f;
    if (m_eventStart && m_everRendered && cudaEventElapsedTime(&ms, (cudaEvent_t)m_eventStart, (cudaEvent_t)m_eventEnd) == cudaSuccess)
      m_duration = ms / 1000.f;
    else
      cudaGetLastError();
    std::memcpy(mem, &m_duration, 4);
    return true;
  }
  if (t == ANARI_INT32 && name == "numSamples" && size >= 4) {
    if (mask & ANARI_WAIT)
      wait();
    std::memcpy(mem, &m_frameID, 4);
    return true;
  }
  if (t == ANARI_BOOL && name == "nextFrameReset" && size >= 4) {
    if (mask & ANARI_WAIT)
      wait();
    if (ready(ANARI_NO_WAIT))
      device->flushCommits();
    checkAccumulationReset();
    const int32_t v = m_nextFrameReset ? 1 : 0;
    std::memcpy(mem, &v, 4);
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------
// Device
// ---------------------------------------------------------------------------------------------------------
static void defaultStatus(const void *, ANARIDevice, ANARIObject, ANARIDataType, ANARIStatusSeverity sev,
    ANARIStatusCode, const char *msg)
{
  if (sev <= ANARI_SEVERITY_WARNING)
    fprintf(stderr, "[visrtx_b200][%s] %s\n",
        sev == ANARI_SEVERITY_FATAL_ERROR ? "FATAL" : (sev == ANARI_SEVERITY_ERROR ? "ERROR" : "WARN "), msg);
}

Device::Device(ANARIStatusCallback cb, const void *userPtr) : Object(nullptr, ANARI_DEVICE, "default")
{
  device = this;
  m_defaultCb = cb ? cb : defaultStatus;
  m_defaultCbUserPtr = userPtr;
  m_cb = m_defaultCb;
  m_cbUserPtr = m_defaultCbUserPtr;
}

Device::~Device()
{
  for (Object *o : m_commitQueue)
    o->refDec(RefType::INTERNAL);
  m_commitQueue.clear();
  for (size_t k = 1; k < m_streams.size(); ++k) {
    GpuScope scope(m_gpus[k]);
    cudaStreamSynchronize((cudaStream_t)m_streams[k]);
    cudaStreamDestroy((cudaStream_t)m_streams[k]);
  }
  if (m_stream) {
    CudaDeviceScope scope(this);
    cudaStreamSynchronize((cudaStream_t)m_stream);
    cudaStreamDestroy((cudaStream_t)m_stream);
  }
  device = nullptr;
}

void Device::message(const Object *src, ANARIStatusSeverity sev, ANARIStatusCode code, const char *msg) const
{
  if (m_cb)
    m_cb(m_cbUserPtr, (ANARIDevice)this, (ANARIObject)src, src ? src->type : ANARI_DEVICE, sev, code, msg);
}

void Device::commitParameters()
{ // VisRTXDevice.cpp:460-473
  m_cb = getParam<ANARIStatusCallback>("statusCallback", ANARI_STATUS_CALLBACK, m_defaultCb);
  m_cbUserPtr = getParam<const void *>("statusCallbackUserData", ANARI_VOID_POINTER, m_defaultCbUserPtr);
  const bool eager = getParam<int32_t>("forceInit", ANARI_BOOL, 0) != 0;
  m_desiredGpuID = getParam<int>("cudaDevice", ANARI_INT32, 0);
  { // "cudaDevices": the GPUs of a multi-GPU device, display GPU first — "0,1,2,3" or an Array1D of INT32
    std::vector<int> gpus;
    const std::string list = getParamString("cudaDevices", "");
    for (size_t i = 0; i < list.size();) {
      size_t j = list.find(',', i);
      if (j == std::string::npos)
        j = list.size();
      if (j > i)
        gpus.push_back(std::atoi(list.substr(i, j - i).c_str()));
      i = j + 1;
    }
    if (Object *a = getParamObject("cudaDevices", ANARI_ARRAY1D)) {
      Array *arr = static_cast<Array *>(a);
      if (arr->elementType == ANARI_INT32 && !arr->onDevice())
        for (size_t i = 0; i < arr->regionSize(); ++i)
          gpus.push_back(((const int32_t *)arr->regionData())[i]);
    }
    if (!gpus.empty())
      m_desiredGpuID = gpus[0];
    if (m_gpuID >= 0 && !gpus.empty() && gpus != m_gpus)
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_OPERATION,
          "visrtx_b200 was already initialized: the new 'cudaDevices' list is ignored.");
    m_desiredGpus = gpus;
    const std::string mode = getParamString("multiGpuMode", "sortLast");
    if (mode != "sortLast" && mode != "sortFirst")
      report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "unknown multiGpuMode '%s' (sortLast | sortFirst)",
          mode.c_str());
    if (m_gpuID < 0)
      m_sortLast = mode != "sortFirst";
  }
  if (m_gpuID >= 0 && m_desiredGpuID != m_gpuID)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_OPERATION,
        "visrtx_b200 was already initialized to use GPU %i: new device number %i is ignored.", m_gpuID,
        m_desiredGpuID);
  if (eager && m_initStatus == 0) {
    report(ANARI_SEVERITY_DEBUG, ANARI_STATUS_NO_ERROR, "eagerly initializing device");
    initDevice();
  }
}

bool Device::initDevice()
{
  if (m_initStatus == 1)
    return true;
  if (m_initStatus == -1) {
    report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNSUPPORTED_DEVICE, "device failed to initialized");
    return false;
  }
  std::lock_guard<std::recursive_mutex> lock(mutex);
  if (m_initStatus != 0)
    return m_initStatus == 1;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_UNSUPPORTED_DEVICE,
        "no CUDA capable devices found (the DVR path has no CPU fallback)");
    m_initStatus = -1;
    return false;
  }
  if (m_desiredGpuID < 0 || m_desiredGpuID >= n) {
    report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_INVALID_ARGUMENT, "cudaDevice %d out of range (%d devices)",
        m_desiredGpuID, n);
    m_initStatus = -1;
    return false;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(m_desiredGpuID);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, m_desiredGpuID);
  if (prop.major < 10)
    report(ANARI_SEVERITY_WARNING, ANARI_STATUS_UNSUPPORTED_DEVICE,
        "GPU %d (%s, sm_%d%d) is not a Blackwell part; the kernels are built for sm_100a only", m_desiredGpuID,
        prop.name, prop.major, prop.minor);
  cudaStream_t s = nullptr;
  const cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    report(ANARI_SEVERITY_FATAL_ERROR, ANARI_STATUS_UNKNOWN_ERROR, "cudaStreamCreate failed: %s",
        cudaGetErrorString(e));
    m_initStatus = -1;
    return false;
  }
  m_stream = s;
  m_gpuID = m_desiredGpuID;
  m_gpus.assign(1, m_gpuID);
  m_streams.assign(1, m_stream);
  // the other GPUs of a multi-GPU device: one private stream each, peer access between every pair (kernels load
  // partial images from and store final pixels into each other's memory over NVLink)
  for (size_t i = 1; i < m_desiredGpus.size(); ++i) {
    const int g = m_desiredGpus[i];
    bool ok = g >= 0 && g < n && std::find(m_gpus.begin(), m_gpus.end(), g) == m_gpus.end();
    cudaStream_t sg = nullptr;
    if (ok) {
      GpuScope scope(g);
      ok = cudaStreamCreateWithFlags(&sg, cudaStreamNonBlocking) == cudaSuccess;
    }
    if (!ok) {
      cudaGetLastError();
      report(ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_ARGUMENT, "cudaDevices: GPU %d is unusable or listed twice; ignored", g);
      continue;
    }
    m_gpus.push_back(g);
    m_streams.push_back(sg);
  }
  for (size_t i = 0; i < m_gpus.size(); ++i)
    for (size_t j = 0; j < m_gpus.size(); ++j) {
      if (i == j)
        continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m_gpus[i], m_gpus[j]);
      GpuScope scope(m_gpus[i]);
      const cudaError_t pe = can ? cudaDeviceEnablePeerAccess(m_gpus[j], 0) : cudaErrorPeerAccessUnsupported;
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        report(ANARI_SEVERITY_ERROR, ANARI_STATUS_UNSUPPORTED_DEVICE,
            "no peer access from GPU %d to GPU %d: multi-GPU rendering disabled", m_gpus[i], m_gpus[j]);
        for (size_t k = 1; k < m_streams.size(); ++k) {
          GpuScope sk(m_gpus[k]);
          cudaStreamDestroy((cudaStream_t)m_streams[k]);
        }
        m_gpus.resize(1);
        m_streams.resize(1);
        i = j = 1u << 20; // leave both loops
      } else
        cudaGetLastError();
    }
  m_initStatus = 1;
  report(ANARI_SEVERITY_DEBUG, ANARI_STATUS_NO_ERROR, "initialised on GPU %d (%s, %d SMs)", m_gpuID, prop.name,
      prop.multiProcessorCount);
  return true;
}

void Device::enqueueCommit(Object *o)
{
  std::lock_guard<std::recursive_mutex> lock(mutex);
  if (std::find(m_commitQueue.begin(), m_commitQueue.end(), o) != m_commitQueue.end())
    return;
  o->refInc(RefType::INTERNAL);
  m_commitQueue.push_back(o);
}

void Device::removeFromQueue(Object *o)
{
  std::lock_guard<std::recursive_mutex> lock(mutex);
  m_commitQueue.erase(std::remove(m_commitQueue.begin(), m_commitQueue.end(), o), m_commitQueue.end());
}

void Device::flushCommits()
{ // helium DeferredCommitBuffer::flush: commitParameters for all, then finalize in dependency order
  std::lock_guard<std::recursive_mutex> lock(mutex);
  int guard = 0;
  while (!m_commitQueue.empty() && guard++ < 16) {
    std::vector<Object *> q;
    q.swap(m_commitQueue);
    std::stable_sort(q.begin(), q.end(), [](Object *a, Object *b) { return a->commitPriority() < b->commitPriority(); });
    for (Object *o : q)
      if (o->useCount(RefType::PUBLIC) + o->useCount(RefType::INTERNAL) > 1) // something besides the queue holds it
        o->commitParameters();
    for (Object *o : q) {
      if (o->useCount(RefType::PUBLIC) + o->useCount(RefType::INTERNAL) > 1) {
        o->finalize();
        o->lastFinalized = newTimeStamp();
        m_lastFinalization = o->lastFinalized;
        o->parametersChanged = false;
        o->notifyObservers(); // volumes observe their field, so a re-uploaded field refreshes its majorants
      }
    }
    for (Object *o : q)
      o->refDec(RefType::INTERNAL);
  }
}

static const char *kExtensions[] = {"ANARI_KHR_CAMERA_ORTHOGRAPHIC", "ANARI_KHR_CAMERA_PERSPECTIVE",
    "ANARI_KHR_FRAME_ACCUMULATION", "ANARI_KHR_FRAME_CHANNEL_PRIMITIVE_ID", "ANARI_KHR_FRAME_CHANNEL_OBJECT_ID",
    "ANARI_KHR_FRAME_CHANNEL_INSTANCE_ID", "ANARI_KHR_FRAME_CHANNEL_ALBEDO", "ANARI_KHR_FRAME_CHANNEL_NORMAL",
    "ANARI_KHR_FRAME_COMPLETION_CALLBACK", "ANARI_KHR_INSTANCE_TRANSFORM", "ANARI_KHR_SPATIAL_FIELD_STRUCTURED_REGULAR",
    "ANARI_KHR_VOLUME_TRANSFER_FUNCTION1D", "ANARI_VISRTX_SPATIAL_FIELD_NANOVDB", "ANARI_KHR_RENDERER_BACKGROUND_COLOR", "ANARI_KHR_DEVICE_SYNCHRONIZATION",
    "ANARI_NV_ARRAY_CUDA", "ANARI_NV_FRAME_BUFFERS_CUDA", "ANARI_VISRTX_CUDA_OUTPUT_BUFFERS", "ANARI_VISRTX_ARRAY_CUDA",
    "ANARI_KHR_CAMERA_DEPTH_OF_FIELD", "ANARI_KHR_RENDERER_AMBIENT_LIGHT", "ANARI_KHR_ARRAY1D_REGION",
    "ANARI_KHR_GEOMETRY_TRIANGLE", "ANARI_KHR_GEOMETRY_SPHERE", "ANARI_KHR_MATERIAL_MATTE", "ANARI_KHR_LIGHT_DIRECTIONAL",
    "ANARI_KHR_LIGHT_POINT",
    "ANARI_VISRTX_B200_DVR", nullptr};

bool Device::getProperty(const std::string &name, ANARIDataType t, void *mem, uint64_t size, uint32_t)
{ // VisRTXDevice.cpp:520-550
  auto wi = [&](int32_t v) {
    if (size < 4)
      return false;
    std::memcpy(mem, &v, 4);
    return true;
  };
  if (t == ANARI_INT32 && name == "version") return wi(DVR_B200_VERSION_MAJOR * 10000 + DVR_B200_VERSION_MINOR * 100);
  if (t == ANARI_INT32 && name == "version.major") return wi(DVR_B200_VERSION_MAJOR);
  if (t == ANARI_INT32 && name == "version.minor") return wi(DVR_B200_VERSION_MINOR);
  if (t == ANARI_INT32 && name == "version.patch") return wi(0);
  if (t == ANARI_STRING_LIST && name == "extension" && size >= sizeof(void *)) {
    const char **p = kExtensions;
    std::memcpy(mem, &p, sizeof(p));
    return true;
  }
  if (t == ANARI_INT32 && name == "cudaDevice") return wi(m_gpuID);
  if (t == ANARI_INT32 && name == "cudaDeviceCount") return wi((int32_t)m_gpus.size()); // GPUs actually in use
  return false;
}

} // namespace b200

// =========================================================================================================
// extern "C" ANARI entry points
// =========================================================================================================
using namespace b200;

struct _ANARILibrary
{
  ANARIStatusCallback cb;
  const void *userPtr;
};

namespace {
inline Device *dev(ANARIDevice d) { return (Device *)d; }
inline Object *obj(ANARIObject o) { return (Object *)o; }

struct ApiScope
{ // serialises API calls on a device and scopes the CUDA device (VisRTXDevice.cpp:790-814)
  explicit ApiScope(Device *d) : lock(d->mutex), cuda(d) {}
  std::lock_guard<std::recursive_mutex> lock;
  CudaDeviceScope cuda;
};

template <typename T, typename... A>
ANARIObject make(ANARIDevice d, A &&...a)
{
  if (!d)
    return nullptr;
  ApiScope s(dev(d));
  return (ANARIObject) new T(dev(d), std::forward<A>(a)...);
}

const char *kDeviceTypes[] = {"default", nullptr};
} // namespace

extern "C" {

ANARILibrary anariLoadLibrary(const char *name, ANARIStatusCallback cb, const void *userPtr)
{
  const std::string n = name ? name : "";
  if (n != "visrtx_b200" && n != "visrtx" && n != "environment") {
    if (cb)
      cb(userPtr, nullptr, nullptr, ANARI_LIBRARY, ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_ARGUMENT,
          "this build only provides the 'visrtx_b200' library (alias 'visrtx')");
    return nullptr;
  }
  return new _ANARILibrary{cb, userPtr};
}
void anariUnloadLibrary(ANARILibrary l) { delete l; }
void anariLoadModule(ANARILibrary, const char *) {}
void anariUnloadModule(ANARILibrary, const char *) {}
const char **anariGetDeviceSubtypes(ANARILibrary) { return kDeviceTypes; }
const char **anariGetDeviceExtensions(ANARILibrary, const char *) { return b200::kExtensions; }

ANARIDevice anariNewDevice(ANARILibrary l, const char *type)
{
  if (!l)
    return nullptr;
  const std::string t = type ? type : "default";
  if (t != "default" && t != "visrtx_b200" && t != "visrtx")
    return nullptr;
  return (ANARIDevice) new Device(l->cb, l->userPtr);
}

ANARIDevice makeVisRTXDevice(ANARIStatusCallback cb, const void *userPtr) { return (ANARIDevice) new Device(cb, userPtr); }

int visrtxGetObjectExtensions(VisRTXExtensions *e, ANARIDevice, ANARIDataType, const char *)
{
  if (!e)
    return 0;
  std::memset(e, 0, sizeof(*e));
  e->VISRTX_ARRAY_CUDA = 1;
  e->VISRTX_CUDA_OUTPUT_BUFFERS = 1;
  e->VISRTX_SPATIAL_FIELD_NANOVDB = 1;
  return 1;
}
int visrtxGetInstanceExtensions(VisRTXExtensions *e, ANARIDevice d, ANARIObject)
{
  return visrtxGetObjectExtensions(e, d, ANARI_UNKNOWN, nullptr);
}

// ---- arrays ----
static ANARIObject newArray(ANARIDevice d, ANARIDataType at, const void *mem, ANARIMemoryDeleter del, const void *ud,
    ANARIDataType et, uint64_t n1, uint64_t n2, uint64_t n3)
{
  if (!d)
    return nullptr;
  ApiScope s(dev(d));
  if (sizeOfType(et) == 0 || n1 * n2 * n3 == 0) {
    dev(d)->report(ANARI_SEVERITY_ERROR, ANARI_STATUS_INVALID_ARGUMENT, "invalid array element type or size");
    return nullptr;
  }
  return (ANARIObject) new Array(dev(d), at, mem, del, ud, et, n1, n2, n3);
}
ANARIArray1D anariNewArray1D(ANARIDevice d, const void *m, ANARIMemoryDeleter del, const void *ud, ANARIDataType t,
    uint64_t n1)
{
  return newArray(d, ANARI_ARRAY1D, m, del, ud, t, n1, 1, 1);
}
ANARIArray2D anariNewArray2D(ANARIDevice d, const void *m, ANARIMemoryDeleter del, const void *ud, ANARIDataType t,
    uint64_t n1, uint64_t n2)
{
  return newArray(d, ANARI_ARRAY2D, m, del, ud, t, n1, n2, 1);
}
ANARIArray3D anariNewArray3D(ANARIDevice d, const void *m, ANARIMemoryDeleter del, const void *ud, ANARIDataType t,
    uint64_t n1, uint64_t n2, uint64_t n3)
{
  return newArray(d, ANARI_ARRAY3D, m, del, ud, t, n1, n2, n3);
}
void *anariMapArray(ANARIDevice d, ANARIArray a)
{
  if (!d || !a)
    return nullptr;
  ApiScope s(dev(d));
  return static_cast<Array *>(obj(a))->map();
}
void anariUnmapArray(ANARIDevice d, ANARIArray a)
{
  if (!d || !a)
    return;
  ApiScope s(dev(d));
  static_cast<Array *>(obj(a))->unmap();
}

// ---- objects ----
ANARILight anariNewLight(ANARIDevice d, const char *t) { return make<Light>(d, std::string(t ? t : "")); }
ANARIGeometry anariNewGeometry(ANARIDevice d, const char *t) { return make<Geometry>(d, std::string(t ? t : "")); }
ANARIMaterial anariNewMaterial(ANARIDevice d, const char *t) { return make<Material>(d, std::string(t ? t : "")); }
ANARISampler anariNewSampler(ANARIDevice d, const char *t) { return make<Object>(d, ANARI_SAMPLER, std::string(t ? t : "")); }
ANARISurface anariNewSurface(ANARIDevice d) { return make<Surface>(d); }
ANARICamera anariNewCamera(ANARIDevice d, const char *t) { return make<Camera>(d, std::string(t ? t : "")); }
ANARISpatialField anariNewSpatialField(ANARIDevice d, const char *t) { return make<SpatialField>(d, std::string(t ? t : "")); }
ANARIVolume anariNewVolume(ANARIDevice d, const char *t) { return make<Volume>(d, std::string(t ? t : "")); }
ANARIGroup anariNewGroup(ANARIDevice d) { return make<Group>(d); }
ANARIInstance anariNewInstance(ANARIDevice d, const char *t) { return make<Instance>(d, std::string(t ? t : "transform")); }
ANARIWorld anariNewWorld(ANARIDevice d) { return make<World>(d); }
ANARIRenderer anariNewRenderer(ANARIDevice d, const char *t) { return make<Renderer>(d, std::string(t ? t : "default")); }
ANARIFrame anariNewFrame(ANARIDevice d) { return make<Frame>(d); }
ANARIObject anariNewObject(ANARIDevice d, const char *objectType, const char *type)
{
  if (d)
    dev(d)->report(ANARI_SEVERITY_WARNING, ANARI_STATUS_INVALID_ARGUMENT, "anariNewObject(%s, %s): no such object",
        objectType ? objectType : "", type ? type : "");
  return nullptr;
}

// ---- parameters / lifetime ----
void anariSetParameter(ANARIDevice d, ANARIObject o, const char *name, ANARIDataType t, const void *mem)
{
  if (!d || !o)
    return;
  ApiScope s(dev(d));
  obj(o)->setParam(name, t, mem);
}
void anariUnsetParameter(ANARIDevice d, ANARIObject o, const char *name)
{
  if (!d || !o)
    return;
  ApiScope s(dev(d));
  obj(o)->unsetParam(name);
}
void anariUnsetAllParameters(ANARIDevice d, ANARIObject o)
{
  if (!d || !o)
    return;
  ApiScope s(dev(d));
  obj(o)->unsetAllParams();
}
static void *mapParamArray(ANARIDevice d, ANARIObject o, const char *name, ANARIDataType at, ANARIDataType et,
    uint64_t n1, uint64_t n2, uint64_t n3, uint64_t *stride)
{
  if (!d || !o)
    return nullptr;
  ApiScope s(dev(d));
  if (sizeOfType(et) == 0 || n1 * n2 * n3 == 0)
    return nullptr;
  Array *a = new Array(dev(d), at, nullptr, nullptr, nullptr, et, n1, n2, n3);
  Object *ao = a;
  obj(o)->setParam(name, at, &ao);
  a->refDec(RefType::PUBLIC);
  if (stride)
    *stride = sizeOfType(et);
  return a->map();
}
void *anariMapParameterArray1D(ANARIDevice d, ANARIObject o, const char *name, ANARIDataType t, uint64_t n1,
    uint64_t *stride)
{
  return mapParamArray(d, o, name, ANARI_ARRAY1D, t, n1, 1, 1, stride);
}
void *anariMapParameterArray2D(ANARIDevice d, ANARIObject o, const char *name, ANARIDataType t, uint64_t n1,
    uint64_t n2, uint64_t *stride)
{
  return mapParamArray(d, o, name, ANARI_ARRAY2D, t, n1, n2, 1, stride);
}
void *anariMapParameterArray3D(ANARIDevice d, ANARIObject o, const char *name, ANARIDataType t, uint64_t n1,
    uint64_t n2, uint64_t n3, uint64_t *stride)
{
  return mapParamArray(d, o, name, ANARI_ARRAY3D, t, n1, n2, n3, stride);
}
void anariUnmapParameterArray(ANARIDevice d, ANARIObject o, const char *name)
{
  if (!d || !o || !name)
    return;
  ApiScope s(dev(d));
  const Param *p = obj(o)->findParam(name);
  if (p && p->object && p->object->type >= ANARI_ARRAY1D && p->object->type <= ANARI_ARRAY3D)
    static_cast<Array *>(p->object)->unmap();
}
void anariCommitParameters(ANARIDevice d, ANARIObject o)
{
  if (!d || !o)
    return;
  ApiScope s(dev(d));
  if (obj(o) == dev(d)) {
    dev(d)->commitParameters();
    return;
  }
  dev(d)->enqueueCommit(obj(o)); // deferred: runs at the next renderFrame / WAIT query (Frame.cu:203)
}
void anariRelease(ANARIDevice d, ANARIObject o)
{
  if (!d || !o)
    return;
  if (obj(o) == dev(d)) {
    Device *dd = dev(d);
    {
      std::lock_guard<std::recursive_mutex> lock(dd->mutex);
      dd->flushCommits();
    }
    dd->refDec(RefType::PUBLIC);
    return;
  }
  ApiScope s(dev(d));
  obj(o)->refDec(RefType::PUBLIC);
}
void anariRetain(ANARIDevice d, ANARIObject o)
{
  if (!d || !o)
    return;
  ApiScope s(dev(d));
  obj(o)->refInc(RefType::PUBLIC);
}

// ---- introspection / properties ----
const char **anariGetObjectSubtypes(ANARIDevice, ANARIDataType t) { return querySubtypes(t); }
const void *anariGetObjectInfo(ANARIDevice, ANARIDataType t, const char *subtype, const char *infoName,
    ANARIDataType infoType)
{
  return queryObjectInfo(t, subtype, infoName, infoType, b200::kExtensions);
}
const void *anariGetParameterInfo(ANARIDevice, ANARIDataType t, const char *subtype, const char *pname,
    ANARIDataType ptype, const char *infoName, ANARIDataType infoType)
{
  return queryParameterInfo(t, subtype, pname, ptype, infoName, infoType);
}
int anariGetProperty(ANARIDevice d, ANARIObject o, const char *name, ANARIDataType t, void *mem, uint64_t size,
    ANARIWaitMask mask)
{
  if (!d || !o || !name || !mem)
    return 0;
  ApiScope s(dev(d));
  if (obj(o) != dev(d) && (mask & ANARI_WAIT) && obj(o)->type != ANARI_FRAME)
    dev(d)->flushCommits();
  return obj(o)->getProperty(name, t, mem, size, mask) ? 1 : 0;
}

// ---- frames ----
const void *anariMapFrame(ANARIDevice d, ANARIFrame f, const char *channel, uint32_t *w, uint32_t *h,
    ANARIDataType *type)
{
  if (!d || !f || !channel)
    return nullptr;
  ApiScope s(dev(d));
  return static_cast<Frame *>(obj(f))->map(channel, w, h, type);
}
void anariUnmapFrame(ANARIDevice, ANARIFrame, const char *) {}
void anariRenderFrame(ANARIDevice d, ANARIFrame f)
{
  if (!d || !f)
    return;
  ApiScope s(dev(d));
  static_cast<Frame *>(obj(f))->renderFrame();
}
int anariFrameReady(ANARIDevice d, ANARIFrame f, ANARIWaitMask m)
{
  if (!d || !f)
    return 0;
  if (m == ANARI_WAIT) { // do not hold the device lock while blocking (other threads may create objects)
    static_cast<Frame *>(obj(f))->wait();
    return 1;
  }
  ApiScope s(dev(d));
  return static_cast<Frame *>(obj(f))->ready(m);
}
void anariDiscardFrame(ANARIDevice, ANARIFrame) {}

} // extern "C"
