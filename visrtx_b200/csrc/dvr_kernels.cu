// dvr_kernels.cu — frame kernel K1 (ray generation + march + background + accumulate/tonemap/encode),
// the sort-last partial kernel, the resolve kernel K3 and the `over` compositing kernel.
// Target: sm_100a only.  See DESIGN.md for the layout / scheduling rationale.
#include <cstdlib>
#include "dvr_internal.h"
#include "dvr_march.cuh"
#include "dvr_dpt.cuh"
#include "dvr_dpt_regen.cuh"
#include "dvr_frame_common.cuh"

namespace dvr {


// ---- cross-GPU flags (DvrPeerSync) ----------------------------------------------------------
__device__ __forceinline__ void signalPeers(const SyncDev &sy)
{
  __threadfence_system(); // everything this GPU wrote for the frame is visible system-wide first
  for (uint32_t i = 0; i < sy.nSignal; ++i)
    *((volatile unsigned int *)sy.signal[i]) = sy.signalValue;
  __threadfence_system();
}

// bounded spin (about 2 s at 1.9 GHz): a missing producer must not hang the GPU
__device__ __forceinline__ bool waitFlags(const unsigned int *flags, uint32_t n, uint32_t value, unsigned int *err)
{
  const long long t0 = clock64();
  for (uint32_t i = 0; i < n; ++i) {
    while ((int)(*((volatile const unsigned int *)&flags[i]) - value) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > 4000000000ll) {
        if (err)
          *err = 1u;
        return false;
      }
    }
  }
  __threadfence_system();
  return true;
}

// retireWarp + signal from the last warp of the grid
__device__ __forceinline__ void retireWarpAndSignal(unsigned int *sched, int lane, const SyncDev &sy)
{
  if (lane == 0) {
    const unsigned int totalWarps = gridDim.x * (blockDim.x >> 5);
    __threadfence();
    const unsigned int done = atomicAdd(&sched[1], 1u);
    if (done == totalWarps - 1u) {
      sched[0] = 0u;
      sched[1] = 0u;
      sched[2] = 0u;
      __threadfence();
      if (sy.nSignal)
        signalPeers(sy);
    }
  }
}

__global__ void dvrSignalFlagsKernel(const __grid_constant__ SyncDev sy) { signalPeers(sy); }

int launchSignalFlags(const SyncDev &sy, cudaStream_t s)
{
  dvrSignalFlagsKernel<<<1, 1, 0, s>>>(sy);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

__global__ void dvrWaitFlagsKernel(const unsigned int *flags, uint32_t n, uint32_t value, unsigned int *err)
{
  waitFlags(flags, n, value, err);
}

int launchWaitFlags(const unsigned int *flags, uint32_t n, uint32_t value, unsigned int *errorFlag, cudaStream_t s)
{
  dvrWaitFlagsKernel<<<1, 1, 0, s>>>(flags, n, value, errorFlag);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}


// ----------------------------------------------------------------------------------------------
// K1: one frame.  Persistent CTAs; every warp pulls 8x4-pixel tiles from a global counter.
// ----------------------------------------------------------------------------------------------
#ifndef DVR_OCC
#define DVR_OCC 2 // minimum resident CTAs per SM the frame kernel is compiled for (register budget)
#endif

// KIND: FIELD_STRUCTURED / FIELD_NANOVDB for the single-volume kernels, -1 for the multi-volume kernel
#ifndef DVR_OCC_NVDB
#define DVR_OCC_NVDB 2 // brick sampling wants registers (no spills at 128), not warps; tree walk alone, batch 1: 3/4/5/6 CTAs = 810/732/642/612 fps
#endif
// G: depth lanes per ray (see marchSegment); the warp's 8x4 tile is then rendered in G passes of 32/G pixels.
#ifndef DVR_DEPTH_LANES
#define DVR_DEPTH_LANES 1
#endif
#ifndef DVR_DEPTH_LANES_NVDB
#define DVR_DEPTH_LANES_NVDB 1
#endif
// K2 (delta tracking) on a structured field is latency- and divergence-bound with 4 warps per scheduler
// (profiles/r02_dpt_ncu.md): 2 / 3 / 4 CTAs per SM = 0.547 / 0.488 / 0.539 ms on C2 (profiles/r02u_brick_staging_probe.md)
#ifndef DVR_OCC_DPT
#define DVR_OCC_DPT 3
#endif
template <bool SKIP, bool STATS, bool SINGLE, int KIND, bool DPT, int G>
__global__ void __launch_bounds__(kBlockThreads,
    (KIND >= FIELD_NANOVDB ? DVR_OCC_NVDB : (DPT && SINGLE && KIND == FIELD_STRUCTURED ? DVR_OCC_DPT : DVR_OCC))) dvrFrameKernel(const __grid_constant__ FrameLaunch P)
{
  static_assert(!(DPT && G != 1) && !(!SINGLE && G != 1), "depth lanes: single-volume marching kernels only");
  __shared__ float4 s_tf[(SINGLE ? 1 : kMaxInlineInstances) * DVR_TF_SIZE];

  const int lane = threadIdx.x & 31;
  const InstanceDev *inst = (P.nInst <= kMaxInlineInstances) ? P.inl : P.ext;
  const int nInst = SINGLE ? 1 : P.nInst;

  // stage the transfer-function tables (4 KiB each) once per CTA
  {
    const int nTab = SINGLE ? 1 : min(P.nInst, kMaxInlineInstances);
    for (int i = threadIdx.x; i < nTab * DVR_TF_SIZE; i += blockDim.x)
      s_tf[i] = __ldg(&inst[i / DVR_TF_SIZE].v.tf[i % DVR_TF_SIZE]);
    __syncthreads();
  }

  MarchStats st{0ull, 0ull};
  unsigned long long raysHit = 0ull;
  const uint32_t nTiles = P.tilesW * P.tilesH;
  const bool centered = P.integrator == DVR_INTEGRATOR_RAYCAST;
  const bool initFrame = P.frameID == 0 && P.checkerboardID <= 0;
  const AccumCtx actx{P.width, P.height, P.format, P.frameID, P.checkerboardID, P.fb};

  for (uint32_t tile = nextTile(P.sched, lane); tile < nTiles; tile = nextTile(P.sched, lane)) {
    // only the tile window is scheduled (the whole launch grid unless the launcher knows the screen rectangle of the
    // volumes; the pixels outside it are then swept by dvrBackgroundSweepKernel on a second stream)
    const uint32_t tyIdx = P.tileY0 + tile / P.tilesW, txIdx = P.tileX0 + tile % P.tilesW;
    if (P.tileRanks > 1u && ((tyIdx / P.tileBand) % P.tileRanks) != P.tileRank)
      continue;
   for (int pass = 0; pass < G; ++pass) {
    // G == 1: lane = pixel of the tile.  G > 1: lane / G = pixel within this pass, lane % G = depth slot.
    const int pix = G == 1 ? lane : pass * (32 / G) + lane / G;
    const uint32_t lx = txIdx * kTileW + (pix % kTileW), ly = tyIdx * kTileH + (pix / kTileW);
    if (lx >= P.launchW || ly >= P.launchH)
      continue;
    uint32_t px = lx, py = ly;
    if (P.checkerboardID >= 0) { // createScreenSample.h:38-46
      px = lx * 2u + (uint32_t)(P.checkerboardID & 1);
      py = ly * 2u + (uint32_t)((P.checkerboardID >> 1) & 1);
    }
    if (px >= P.width || py >= P.height)
      continue;

    if (!DPT && P.missValid
        && ((int)px < P.missX0 || (int)px >= P.missX1 || (int)py < P.missY0 || (int)py >= P.missY1)) {
      // No ray of this pixel can enter a volume.  Launches that need the ray direction (normal channel) never set
      // missValid.
      shadeMissedPixel(actx, px, py, P.background, P.bgTex, centered, P.invW, P.invH, P.numIterations, initFrame);
      continue;
    }
    Philox rng;
    rng.init((unsigned long long)(int)(py * P.width + px), (unsigned long long)P.frameID * 512ull);
    DptPath path{0, f3(1.f, 1.f, 1.f)}; // PathData lives outside the iteration loop in the reference

    for (int it = 0; it < P.numIterations; ++it) {
      // makePrimaryRay, cameraCreateRay.h:74-81
      const float4 r = rng.uniform4();
      const float sx = __fmul_rn(centered ? (float)px : __fadd_rn((float)px, r.x), P.invW);
      const float sy = __fmul_rn(centered ? (float)py : __fadd_rn((float)py, r.y), P.invH);
      float3 org, dir;
      cameraCreateRay(P.cam, sx, sy, r.z, r.w, org, dir);

      if (DPT && P.integrator == DVR_INTEGRATOR_TEST) { // Test_ptx.cu:52-69: one sample, no scene access
        accumResults(actx, px, py, make_float4(dir.x, dir.y, dir.z, 1.f), 1.f, dir, f3(-dir.x, -dir.y, -dir.z), ~0u,
            ~0u, ~0u, 0, initFrame);
        break;
      }
      if (DPT) {
        // DiffusePathTracer_ptx.cu:96-215: colour = Lw * ambient (or the background when nothing scattered),
        // alpha 1; depth / ids are never set by the reference's loop (its `depth == 0` test runs after the
        // increment), so depth stays tmax and the ids ~0u; albedo channel = background, normal = primary dir
        float3 c;
        const float4 bgd = backgroundAt(P.bgTex, P.background, sx, sy);
        if (SINGLE)
          c = dptTracePath<true, KIND>(P.inl, 1, TfSelectSingle{s_tf}, org, dir, P.maxDepth, P.occlusionDistance,
              P.ambientIntensity, bgd, rng, path);
        else
          c = dptTracePath<false, -1>(inst, nInst, TfSelectShared{s_tf, inst}, org, dir, P.maxDepth,
              P.occlusionDistance, P.ambientIntensity, bgd, rng, path);
        accumResults(actx, px, py, make_float4(c.x, c.y, c.z, 1.f), FLT_MAX, f3(bgd.x, bgd.y, bgd.z), dir, ~0u, ~0u,
            ~0u, it, initFrame && it == 0);
        continue;
      }

      float3 color = f3(0.f, 0.f, 0.f);
      float opacity = 0.f;
      uint32_t objID = ~0u, instID = ~0u;
      bool anyHit = false;
      float volumeDepth;
      if (SINGLE)
        volumeDepth = rayMarchAllVolumes<SKIP, false, STATS, true, KIND, G>(P.inl, 1, TfSelectSingle{s_tf}, org, dir,
            FLT_MAX, P.invSamplingRate, rng, color, opacity, objID, instID, st, P.cellBitmap, anyHit);
      else
        volumeDepth = rayMarchAllVolumes<SKIP, false, STATS, false, -1>(inst, nInst, TfSelectShared{s_tf, inst}, org, dir,
            FLT_MAX, P.invSamplingRate, rng, color, opacity, objID, instID, st, P.cellBitmap, anyHit);
      const bool writer = G == 1 || (lane & (G - 1)) == 0; // every depth lane holds the same result; one stores it
      if (STATS && anyHit && writer)
        raysHit++;

      // Raycast_ptx.cu:139-166 (no-surface branch)
      const float depth = fminf(1e30f, volumeDepth);
      color = color * opacity;
      const float4 bg = backgroundAt(P.bgTex, P.background, sx, sy);
      const float oneMinus = __fsub_rn(1.f, opacity);
      color.x = __fmaf_rn(bg.x, oneMinus, color.x);
      color.y = __fmaf_rn(bg.y, oneMinus, color.y);
      color.z = __fmaf_rn(bg.z, oneMinus, color.z);
      opacity = __fmaf_rn(bg.w, oneMinus, opacity);
      // outputColor/outputOpacity start at 0: accumulateValue(out, c, 0) == c
      if (writer)
        accumResults(actx, px, py, make_float4(color.x, color.y, color.z, opacity), depth, color, dir, 0u, objID,
            instID, it, initFrame && it == 0);
    }
   } // pass
  }

  if (STATS) {
    const unsigned long long a = warpSum(st.taken), b = warpSum(st.skipped), c = warpSum(raysHit);
    if (lane == 0 && P.stats) {
      atomicAdd(&P.stats->samplesTaken, a);
      atomicAdd(&P.stats->samplesSkipped, b);
      atomicAdd(&P.stats->raysHit, c);
    }
  }
  retireWarp(P.sched, lane);
}

// Background sweep: every pixel outside the tile-aligned screen rectangle of the volumes gets what a missed ray
// produces (see the in-kernel fast path above), one thread per pixel in row-major order — fully coalesced 512 B
// accumulation and 128 B colour / mirror stores.  Runs on a second stream next to the frame kernel, which then
// schedules only the tiles inside the rectangle.
__global__ void __launch_bounds__(256) dvrBackgroundSweepKernel(const __grid_constant__ FrameLaunch P)
{
  const bool initFrame = P.frameID == 0 && P.checkerboardID <= 0;
  const bool centered = P.integrator == DVR_INTEGRATOR_RAYCAST;
  const AccumCtx actx{P.width, P.height, P.format, P.frameID, P.checkerboardID, P.fb};
  const size_t n = (size_t)P.width * P.height;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / P.width), x = (uint32_t)(i - (size_t)y * P.width);
    if ((int)y >= P.missY0 && (int)y < P.missY1 && (int)x >= P.missX0 && (int)x < P.missX1)
      continue; // the tiles own the inside
    shadeMissedPixel(actx, x, y, P.background, P.bgTex, centered, P.invW, P.invH, P.numIterations, initFrame);
  }
}

int launchBackgroundSweep(const FrameLaunch &p, cudaStream_t s)
{
  const size_t n = (size_t)p.width * p.height;
  if (p.fb.outMirror) {
    // Colour is mirrored to pinned host memory: the sweep is bound by PCIe, not by the SMs.  A thin grid of small
    // CTAs (64 threads, ~3 K registers) fits beside the frame kernel's two resident CTAs per SM, so its posted
    // stores drain over the whole march instead of holding a CTA slot of the march hostage.
    dvrBackgroundSweepKernel<<<(unsigned)smCount() * 2u, 64, 0, s>>>(p);
  } else {
    // device-only: a wide grid that is done in a few tens of microseconds and then frees the SMs
    const unsigned want = (unsigned)((n + 255) / 256);
    const unsigned cap = (unsigned)smCount() * 2u;
    dvrBackgroundSweepKernel<<<want < cap ? want : cap, 256, 0, s>>>(p);
  }
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

template <bool SKIP, bool STATS, bool SINGLE, int KIND, bool DPT>
static int launchFrameT(const FrameLaunch &p, cudaStream_t s)
{
  constexpr int G = (DPT || !SINGLE) ? 1 : (KIND >= FIELD_NANOVDB ? DVR_DEPTH_LANES_NVDB : DVR_DEPTH_LANES);
  static int blocksPerSm = 0;
  if (blocksPerSm == 0) {
    DVR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &blocksPerSm, dvrFrameKernel<SKIP, STATS, SINGLE, KIND, DPT, G>, kBlockThreads, 0));
    if (blocksPerSm < 1)
      blocksPerSm = 1;
  }
  const uint32_t nTiles = p.tilesW * p.tilesH;
  const uint32_t warpsPerBlock = kBlockThreads / 32;
  uint32_t grid = (uint32_t)(smCount() * blocksPerSm);
  const uint32_t need = (nTiles + warpsPerBlock - 1) / warpsPerBlock;
  if (grid > need)
    grid = need;
  if (grid == 0)
    grid = 1;
  dvrFrameKernel<SKIP, STATS, SINGLE, KIND, DPT, G><<<grid, kBlockThreads, 0, s>>>(p);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

template <bool SKIP, bool STATS>
static int launchFrameK(const FrameLaunch &p, cudaStream_t s)
{
  if (p.nInst != 1)
    return launchFrameT<SKIP, STATS, false, -1, false>(p, s);
  if (p.inl[0].v.f.kind == FIELD_NANOVDB)
    return launchFrameT<SKIP, STATS, true, FIELD_NANOVDB, false>(p, s);
  if (p.inl[0].v.f.kind == FIELD_NANOVDB_QUANT) {
    if (STATS) // instrumentation is not worth a dedicated instantiation: the multi-volume kernel dispatches at run time
      return launchFrameT<SKIP, STATS, false, -1, false>(p, s);
    return launchFrameT<SKIP, false, true, FIELD_NANOVDB_QUANT, false>(p, s);
  }
  return launchFrameT<SKIP, STATS, true, FIELD_STRUCTURED, false>(p, s);
}

// K2r: persistent lanes with pixel regeneration (dvr_dpt_regen.cuh)
template <int KIND>
static int launchDptRegenT(const FrameLaunch &p, cudaStream_t s)
{
  static int blocksPerSm = 0;
  if (blocksPerSm == 0) {
    DVR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, dvrDptRegenKernel<KIND>, kBlockThreads, 0));
    if (blocksPerSm < 1)
      blocksPerSm = 1;
  }
  const uint32_t nPix = p.tilesW * p.tilesH * 32u;
  uint32_t grid = (uint32_t)(smCount() * blocksPerSm);
  const uint32_t need = (nPix + kBlockThreads - 1) / kBlockThreads;
  if (grid > need)
    grid = need;
  if (grid == 0)
    grid = 1;
  dvrDptRegenKernel<KIND><<<grid, kBlockThreads, 0, s>>>(p);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

static bool dptRegenEnabled()
{
  // Opt-in ("1"): measured SLOWER than the tile kernel on the C2 dpt scene at every refill threshold (0.55-0.90 ms
  // against 0.488 ms, profiles/r02_dpt_ncu.md), so the tile kernel stays the default.  Read at every launch so that
  // one process can A/B both (bit-identity test).
  const char *e = std::getenv("DVR_B200_DPT_REGEN");
  return e && e[0] == '1';
}

int launchFrame(const FrameLaunch &p, bool skip, bool stats, cudaStream_t s)
{
  if (p.integrator == DVR_INTEGRATOR_DPT && p.nInst == 1 && p.inl[0].identity && dptRegenEnabled()) {
    if (p.inl[0].v.f.kind == FIELD_NANOVDB)
      return launchDptRegenT<FIELD_NANOVDB>(p, s);
    if (p.inl[0].v.f.kind == FIELD_NANOVDB_QUANT)
      return launchDptRegenT<FIELD_NANOVDB_QUANT>(p, s);
    return launchDptRegenT<FIELD_STRUCTURED>(p, s);
  }
  if (p.integrator == DVR_INTEGRATOR_DPT || p.integrator == DVR_INTEGRATOR_TEST) {
    // delta tracking has no fixed-step lattice (no SKIP/STATS variants); the test renderer shares the instantiation
    if (p.nInst != 1)
      return launchFrameT<false, false, false, -1, true>(p, s);
    if (p.inl[0].v.f.kind == FIELD_NANOVDB)
      return launchFrameT<false, false, true, FIELD_NANOVDB, true>(p, s);
    if (p.inl[0].v.f.kind == FIELD_NANOVDB_QUANT)
      return launchFrameT<false, false, true, FIELD_NANOVDB_QUANT, true>(p, s);
    return launchFrameT<false, false, true, FIELD_STRUCTURED, true>(p, s);
  }
  if (stats)
    return skip ? launchFrameK<true, true>(p, s) : launchFrameK<false, true>(p, s);
  return skip ? launchFrameK<true, false>(p, s) : launchFrameK<false, false>(p, s);
}

// ----------------------------------------------------------------------------------------------
// self-test of latticeAdvance(): closed form vs the literal `while (n > 0 && t <= tUpper) t += step` loop on
// pseudo-random operands (Philox), including power-of-two steps (round-to-even ties) and tiny / huge t
__global__ void dvrSelftestLatticeKernel(uint32_t count, unsigned long long seed, unsigned int *mismatches)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count)
    return;
  Philox rng;
  rng.init(seed + i, 0ull);
  const float4 r = rng.uniform4();
  const float4 q = rng.uniform4();
  float t, step;
  switch (i % 5u) {
  case 0: t = r.x * 4000.f; step = r.y * 2.f + 1e-3f; break;
  case 1: t = r.x * 10.f; step = exp2f(floorf(r.y * 15.f) - 12.f); break;
  case 2: t = exp2f(floorf(r.x * 17.f) - 3.f) * (1.f + r.z); step = exp2f(floorf(r.y * 24.f) - 20.f) * 1.5f; break;
  case 3: t = r.x * 1e-3f; step = r.y * 0.1f; break;
  default: t = r.x * 3000.f + 1000.f; step = 1.f; break;
  }
  const int n = 1 + (int)(q.x * 3000.f);
  const float tUpper = t + q.y * step * (float)n * 1.5f;
  float a = t;
  int ka = 0;
  for (int k = n; k > 0 && a <= tUpper; --k) {
    a = __fadd_rn(a, step);
    ++ka;
  }
  int kb;
  const float b = latticeAdvance(t, step, n, tUpper, kb);
  if (__float_as_int(a) != __float_as_int(b) || ka != kb)
    atomicAdd(mismatches, 1u);
}

int launchSelftestLattice(uint32_t count, unsigned long long seed, unsigned int *mismatches, cudaStream_t s)
{
  dvrSelftestLatticeKernel<<<(count + 255) / 256, 256, 0, s>>>(count, seed, mismatches);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// ----------------------------------------------------------------------------------------------
// sort-last partial render: premultiplied (C,A) + entry depth of ONE slab on the global lattice
// ----------------------------------------------------------------------------------------------
template <bool SKIP, bool STATS>
__global__ void __launch_bounds__(kBlockThreads, 2) dvrPartialKernel(const __grid_constant__ PartialLaunch P)
{
  __shared__ float4 s_tf[DVR_TF_SIZE];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < DVR_TF_SIZE; i += blockDim.x)
    s_tf[i] = __ldg(&P.inst.v.tf[i]);
  __syncthreads();

  MarchStats st{0ull, 0ull};
  const uint32_t nTiles = P.tilesW * P.tilesH;
  const bool centered = P.integrator == DVR_INTEGRATOR_RAYCAST;
  for (uint32_t tile = nextTile(P.sched, lane); tile < nTiles; tile = nextTile(P.sched, lane)) {
    const uint32_t tyIdx = P.tileY0 + tile / P.tilesW, txIdx = P.tileX0 + tile % P.tilesW;
    const uint32_t px = txIdx * kTileW + (lane % kTileW), py = tyIdx * kTileH + (lane / kTileW);
    if (px >= P.width || py >= P.height)
      continue;
    Philox rng;
    rng.init((unsigned long long)(int)(py * P.width + px), (unsigned long long)P.frameID * 512ull);
    const float4 r = rng.uniform4();
    const float sx = __fmul_rn(centered ? (float)px : __fadd_rn((float)px, r.x), P.invW);
    const float sy = __fmul_rn(centered ? (float)py : __fadd_rn((float)py, r.y), P.invH);
    float3 org, dir;
    cameraCreateRay(P.cam, sx, sy, r.z, r.w, org, dir);
    float3 color = f3(0.f, 0.f, 0.f);
    float opacity = 0.f;
    uint32_t objID = ~0u, instID = ~0u;
    bool anyHit = false;
    const float depth = rayMarchAllVolumes<SKIP, true, STATS, true, FIELD_STRUCTURED>(&P.inst, 1, TfSelectSingle{s_tf}, org, dir, FLT_MAX,
        P.invSamplingRate, rng, color, opacity, objID, instID, st, P.cellBitmap, anyHit);
    const uint32_t idx = px + py * P.width;
    P.partialRgba[idx] = make_float4(color.x, color.y, color.z, opacity);
    P.partialDepth[idx] = fminf(1e30f, depth);
  }
  if (STATS) {
    const unsigned long long a = warpSum(st.taken), b = warpSum(st.skipped);
    if (lane == 0 && P.stats) {
      atomicAdd(&P.stats->samplesTaken, a);
      atomicAdd(&P.stats->samplesSkipped, b);
    }
  }
  retireWarpAndSignal(P.sched, lane, P.sync);
}

template <bool SKIP, bool STATS>
static int launchPartialT(const PartialLaunch &p, cudaStream_t s)
{
  static int bps = 0;
  if (bps == 0) {
    DVR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, dvrPartialKernel<SKIP, STATS>, kBlockThreads, 0));
    if (bps < 1)
      bps = 1;
  }
  const uint32_t nTiles = p.tilesW * p.tilesH;
  uint32_t grid = (uint32_t)(smCount() * bps);
  const uint32_t need = (nTiles + 7) / 8;
  if (grid > need)
    grid = need;
  if (grid == 0)
    grid = 1;
  dvrPartialKernel<SKIP, STATS><<<grid, kBlockThreads, 0, s>>>(p);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

int launchPartial(const PartialLaunch &p, cudaStream_t s)
{
  if (p.stats)
    return p.skip ? launchPartialT<true, true>(p, s) : launchPartialT<false, true>(p, s);
  return p.skip ? launchPartialT<true, false>(p, s) : launchPartialT<false, false>(p, s);
}

// ----------------------------------------------------------------------------------------------
// `over` compositing of two partial images (front-to-back, premultiplied)
// ----------------------------------------------------------------------------------------------
__global__ void dvrCompositeOverKernel(float4 *__restrict__ front, float *__restrict__ frontDepth,
    const float4 *__restrict__ back, const float *__restrict__ backDepth, size_t begin, size_t end,
    bool backIsInFront)
{
  const size_t i = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= end)
    return;
  float4 a = front[i];
  float4 b = back[i];
  if (backIsInFront) {
    const float4 t = a;
    a = b;
    b = t;
  }
  const float k = 1.f - a.w;
  front[i] = make_float4(a.x + k * b.x, a.y + k * b.y, a.z + k * b.z, a.w + k * b.w);
  if (frontDepth && backDepth)
    frontDepth[i] = fminf(frontDepth[i], backDepth[i]);
}

int launchCompositeOver(float4 *front, float *frontDepth, const float4 *back, const float *backDepth,
    size_t begin, size_t end, bool backIsInFront, cudaStream_t s)
{
  if (end <= begin)
    return DVR_OK;
  const size_t n = end - begin;
  dvrCompositeOverKernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      front, frontDepth, back, backDepth, begin, end, backIsInFront);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// ----------------------------------------------------------------------------------------------
// K3: resolve a composited partial image (quirk Q2 + background + accumulate/tonemap/encode)
// ----------------------------------------------------------------------------------------------
// background of the (single) pixel-sample of pixel i of a resolve launch: the constant colour, or the image at the
// screen coordinate makePrimaryRay gave that sample (first four draws of the pixel's Philox stream)
__device__ __forceinline__ float4 resolveBackground(const ResolveLaunch &R, uint32_t px, uint32_t py)
{
  if (!R.bgTex)
    return R.background;
  float sx = (float)px, sy = (float)py;
  if (!R.centered) {
    Philox rng;
    rng.init((unsigned long long)(int)(py * R.width + px), (unsigned long long)R.frameID * 512ull);
    const float4 r = rng.uniform4();
    sx = __fadd_rn(sx, r.x);
    sy = __fadd_rn(sy, r.y);
  }
  return tex2D<float4>(R.bgTex, __fmul_rn(sx, R.invW), __fmul_rn(sy, R.invH));
}

__device__ __forceinline__ void resolvePixel(const ResolveLaunch &R, size_t i, float4 pc, float pd, float4 bg)
{
  float3 color = f3(pc.x, pc.y, pc.z);
  float opacity = pc.w;
  color = color * opacity; // Raycast_ptx.cu:159
  const float oneMinus = __fsub_rn(1.f, opacity);
  color.x = __fmaf_rn(bg.x, oneMinus, color.x);
  color.y = __fmaf_rn(bg.y, oneMinus, color.y);
  color.z = __fmaf_rn(bg.z, oneMinus, color.z);
  opacity = __fmaf_rn(bg.w, oneMinus, opacity);
  const AccumCtx P{R.width, R.height, R.format, R.frameID, -1, R.fb};
  const bool hit = pd < 1e30f;
  const uint32_t px = (uint32_t)(i % R.width), py = (uint32_t)(i / R.width);
  accumResults(P, px, py, make_float4(color.x, color.y, color.z, opacity), pd, color, f3(0.f, 0.f, 0.f), 0u,
      hit ? R.objId : ~0u, hit ? R.instId : ~0u, 0, R.frameID == 0);
}

__global__ void dvrResolveKernel(const __grid_constant__ ResolveLaunch R)
{
  const size_t i = R.pixelBegin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= R.pixelEnd)
    return;
  const float4 pc = R.partialRgba[i];
  const float pd = R.partialDepth ? R.partialDepth[i] : 1e30f;
  resolvePixel(R, i, pc, pd, resolveBackground(R, (uint32_t)(i % R.width), (uint32_t)(i / R.width)));
}

int launchResolve(const ResolveLaunch &r, cudaStream_t s)
{
  if (r.pixelEnd <= r.pixelBegin)
    return DVR_OK;
  const size_t n = r.pixelEnd - r.pixelBegin;
  dvrResolveKernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(r);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// ----------------------------------------------------------------------------------------------
// sort-last direct send: composite the slabs' partial images (peer loads over NVLink) in per-pixel
// view order and resolve, in one kernel
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dvrPeerResolveKernel(const __grid_constant__ PeerResolveLaunch L)
{
  const size_t i = L.r.pixelBegin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < L.r.pixelEnd) {
    const uint32_t px = (uint32_t)(i % L.r.width), py = (uint32_t)(i / L.r.width);
    bool hit = true;
    bool ascending = true;
    float4 bg = L.r.background;
    if (L.cull && L.missValid
        && ((int)px < L.missX0 || (int)px >= L.missX1 || (int)py < L.missY0 || (int)py >= L.missY1)) {
      hit = false; // outside the screen rectangle of the bounds: no ray of this pixel can hit
      bg = resolveBackground(L.r, px, py);
    } else {
      // the primary ray exactly as the partial march generated it (same Philox stream, same arithmetic)
      Philox rng;
      rng.init((unsigned long long)(int)(py * L.r.width + px), (unsigned long long)L.r.frameID * 512ull);
      const float4 r = rng.uniform4();
      const bool centered = L.integrator == DVR_INTEGRATOR_RAYCAST;
      const float sx = __fmul_rn(centered ? (float)px : __fadd_rn((float)px, r.x), L.invW);
      const float sy = __fmul_rn(centered ? (float)py : __fadd_rn((float)py, r.y), L.invH);
      float3 org, dir;
      cameraCreateRay(L.cam, sx, sy, r.z, r.w, org, dir);
      bg = backgroundAt(L.r.bgTex, L.r.background, sx, sy);
      float3 lo = org, ld = dir;
      if (!L.identity) {
        lo = xfmPoint(L.xfm, org);
        ld = xfmVector(L.xfm, dir);
      }
      if (L.cull) {
        float t0, t1;
        hit = intersectVolumeBox(L.boundsLo, L.boundsHi, lo, ld, 0.f, FLT_MAX, t0, t1);
      }
      ascending = ld.z >= 0.f; // rays travelling towards +z (object space) meet the low-z slab first
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float depth = 1e30f;
    if (hit) {
      // (b) issue all peer loads first (independent), then composite front to back
      float4 part[kMaxSlabs];
      float pdep[kMaxSlabs];
#pragma unroll
      for (int k = 0; k < kMaxSlabs; ++k) {
        if (k < L.nSlabs) {
          const int sidx = ascending ? k : L.nSlabs - 1 - k;
          part[k] = __ldcv(&L.rgba[sidx][i]); // volatile load: another GPU wrote this line
          pdep[k] = L.depth[sidx] ? __ldcv(&L.depth[sidx][i]) : 1e30f;
        }
      }
#pragma unroll
      for (int k = 0; k < kMaxSlabs; ++k) {
        if (k < L.nSlabs && acc.w < 0.99f) { // the one-pass march takes no sample once opacity >= 0.99
          const float w = __fsub_rn(1.f, acc.w);
          acc.x = __fmaf_rn(w, part[k].x, acc.x);
          acc.y = __fmaf_rn(w, part[k].y, acc.y);
          acc.z = __fmaf_rn(w, part[k].z, acc.z);
          acc.w = __fmaf_rn(w, part[k].w, acc.w);
          depth = fminf(depth, pdep[k]);
        }
      }
    }
    resolvePixel(L.r, i, acc, depth, bg);
  }
}

int launchPeerResolve(const PeerResolveLaunch &p, cudaStream_t s)
{
  // The cross-GPU ordering of the sync variant brackets the launch with two one-thread kernels: a single
  // system-scope wait before (instead of one fence per CTA) and a single release after (the kernel
  // boundary orders every peer store of the composite before the flag).
  if (p.sync.nWait) {
    const int rc = launchWaitFlags(p.sync.wait, p.sync.nWait, p.sync.waitValue, p.sync.errorFlag, s);
    if (rc != DVR_OK)
      return rc;
  }
  if (p.r.pixelEnd > p.r.pixelBegin) {
    const size_t n = p.r.pixelEnd - p.r.pixelBegin;
    dvrPeerResolveKernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    DVR_CUDA(cudaGetLastError());
    countLaunch();
  }
  if (p.sync.nSignal)
    return launchSignalFlags(p.sync, s);
  return DVR_OK;
}

// ----------------------------------------------------------------------------------------------
// Sort-last frame, fused: march + exchange + composite + resolve in one launch per GPU.
//
//   phase 1  every warp pulls 8x4 tiles of the screen window in the SAME order on every GPU and marches its slab into
//            its partial image.  Tiles are grouped into regions of `tilesPerRegion` consecutive tiles; the warp that
//            completes a region's last tile publishes "region r of frame seq done on rank me" into the table of the
//            region's owner (r mod N) with a system-scope release.
//   phase 2  once the tile queue is empty, warps pull (a) chunks of this rank's share of the pixels outside the window
//            (background only, no peer data) and (b) the tiles of the regions this rank owns, in region order: wait
//            until all N ranks flagged the region, load the N partial pixels over NVLink (volatile peer loads), `over`
//            them in per-pixel view order, resolve (Q2, background, accumulate, tonemap, encode) and store into the
//            display GPU's frame.  Regions complete in order over the frame time, so at the end of the march only
//            the last regions' composites remain: the exchange hides behind the tail of the long rays.
//   retire   the last warp re-arms the counters and publishes "rank me resolved frame seq" to every rank (partial
//            buffers alternate: nobody overwrites a buffer before its readers of two frames ago have finished); the
//            display rank does not retire before it holds every rank's flag.
// Replaces dvrPartialKernel + dvrWaitFlagsKernel + dvrPeerResolveKernel + dvrSignalFlagsKernel (3 serialized
// launches after the march: 42 us of a 180 us step on 8 GPUs).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globalTimerNs()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Flag protocol of the fused frame.  Everything a consumer reads after a flag is either homed in the producer's own
// memory (partial pixels: stored, fenced, THEN flagged — a load issued after the flag was seen finds them in the
// producer's L2) or fetched with strong system-scope loads that bypass L1, so consumers need no fence of their own:
// acquire-flavoured operations compile to CCTL.IVALL, which would wipe the L1 / texture cache under the warps that are
// still marching on the same SM, and a fence.sc.sys per composite item costs microseconds once peer stores are in
// flight (measured: 24 us of composite tail at N = 2 with it, profiles/r02_sort_last_fused.md).
__device__ __forceinline__ bool spinUntil(const unsigned int *flag, uint32_t value, unsigned int *err,
    unsigned sleepNs = 100u)
{
  const long long t0 = clock64();
  while ((int)(*((volatile const unsigned int *)flag) - value) < 0) {
    __nanosleep(sleepNs);
    if (clock64() - t0 > 4000000000ll) { // ~2 s: a missing producer must not hang the GPU
      if (err)
        *err = 1u;
      return false;
    }
  }
  return true;
}

// composite + resolve of pixel (px, py) from the N partial images, exactly dvrPeerResolveKernel's arithmetic
__device__ __forceinline__ void compositePixel(const PeerResolveLaunch &L, uint32_t px, uint32_t py, bool insideWindow)
{
  const size_t i = (size_t)py * L.r.width + px;
  bool hit = insideWindow;
  bool ascending = true;
  float4 bg = L.r.background;
  if (!insideWindow) {
    bg = resolveBackground(L.r, px, py);
  } else {
    Philox rng;
    rng.init((unsigned long long)(int)(py * L.r.width + px), (unsigned long long)L.r.frameID * 512ull);
    const float4 r = rng.uniform4();
    const bool centered = L.integrator == DVR_INTEGRATOR_RAYCAST;
    const float sx = __fmul_rn(centered ? (float)px : __fadd_rn((float)px, r.x), L.invW);
    const float sy = __fmul_rn(centered ? (float)py : __fadd_rn((float)py, r.y), L.invH);
    float3 org, dir;
    cameraCreateRay(L.cam, sx, sy, r.z, r.w, org, dir);
    bg = backgroundAt(L.r.bgTex, L.r.background, sx, sy);
    float3 lo = org, ld = dir;
    if (!L.identity) {
      lo = xfmPoint(L.xfm, org);
      ld = xfmVector(L.xfm, dir);
    }
    float t0, t1;
    hit = intersectVolumeBox(L.boundsLo, L.boundsHi, lo, ld, 0.f, FLT_MAX, t0, t1);
    ascending = ld.z >= 0.f;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float depth = 1e30f;
  if (hit) {
    // two half-batches of peer loads in flight (register budget shared with the march)
#pragma unroll
    for (int h = 0; h < kMaxSlabs; h += 8) {
      if (h < L.nSlabs) {
        float4 part[8];
        float pdep[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (h + k < L.nSlabs) {
            const int sidx = ascending ? h + k : L.nSlabs - 1 - (h + k);
            part[k] = __ldcv(&L.rgba[sidx][i]);
            pdep[k] = __ldcv(&L.depth[sidx][i]);
          }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (h + k < L.nSlabs && acc.w < 0.99f) { // the one-pass march takes no sample once opacity >= 0.99
            const float w = __fsub_rn(1.f, acc.w);
            acc.x = __fmaf_rn(w, part[k].x, acc.x);
            acc.y = __fmaf_rn(w, part[k].y, acc.y);
            acc.z = __fmaf_rn(w, part[k].z, acc.z);
            acc.w = __fmaf_rn(w, part[k].w, acc.w);
            depth = fminf(depth, pdep[k]);
          }
      }
    }
  }
  resolvePixel(L.r, i, acc, depth, bg);
}

// One chunk (256 pixels) of this rank's share of the pixels outside the window.  Constant background, frames after
// the first: every such pixel has held the same accumulation value since the reset (same formula, no per-pixel
// input), so the warp derives the new value ONCE from its first pixel — the same float operations accumResults would
// do per pixel, bit for bit — and the chunk becomes a pure fill of accumulation + colour (depth / ids do not change:
// 1e30 < 1e30 is false).  Frame 0 and image backgrounds take the per-pixel path.
__device__ __forceinline__ void slabBackgroundChunk(const SlabFrameLaunch &S, uint32_t chunk, int lane)
{
  const PartialLaunch &P = S.m;
  const ResolveLaunch &R = S.c.r;
  const int wx0 = (int)(P.tileX0 * kTileW), wy0 = (int)(P.tileY0 * kTileH);
  const int wx1 = (int)((P.tileX0 + P.tilesW) * kTileW), wy1 = (int)((P.tileY0 + P.tilesH) * kTileH);
  const size_t base = S.bgPixelBegin + (size_t)chunk * 256;
  const bool uniform = R.frameID > 0 && !R.bgTex && !R.fb.albedo && !R.fb.normal;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), outF = acc;
  uint32_t outU = 0u;
  bool have = false;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const size_t i = base + (size_t)k * 32 + lane;
    const bool inRange = i < S.bgPixelEnd;
    // (frames have fewer than 2^32 pixels: a 32-bit division instead of the emulated 64-bit one, 8 per chunk)
    const uint32_t i32 = (uint32_t)i;
    const uint32_t py = inRange ? i32 / P.width : 0u, px = inRange ? i32 - py * P.width : 0u;
    const bool mine = inRange && !((int)px >= wx0 && (int)px < wx1 && (int)py >= wy0 && (int)py < wy1);
    if (!uniform) {
      if (mine)
        compositePixel(S.c, px, py, false);
      continue;
    }
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (m == 0u)
      continue;
    if (!have) { // first pixel of the chunk that is ours: its accumulation value stands for all of them
      const int src = __ffs(m) - 1;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane == src)
        a = __ldcg(&R.fb.accum[i]);
      a.x = __shfl_sync(0xffffffffu, a.x, src);
      a.y = __shfl_sync(0xffffffffu, a.y, src);
      a.z = __shfl_sync(0xffffffffu, a.z, src);
      a.w = __shfl_sync(0xffffffffu, a.w, src);
      // accumResults for colour = background (0 * 0 + bg * (1 - 0)): tonemap, add, average, inverse tonemap, encode
      const float4 bg = R.background;
      const float tmDen = __fadd_rn(1.0f, fmaxf(0.0f, fmaxf(fmaxf(bg.x, bg.y), bg.z)));
      acc = make_float4(__fadd_rn(a.x, __fdiv_rn(bg.x, tmDen)), __fadd_rn(a.y, __fdiv_rn(bg.y, tmDen)),
          __fadd_rn(a.z, __fdiv_rn(bg.z, tmDen)), __fadd_rn(a.w, bg.w));
      const float div = float(R.frameID + 1);
      float4 c = make_float4(__fdiv_rn(acc.x, div), __fdiv_rn(acc.y, div), __fdiv_rn(acc.z, div), __fdiv_rn(acc.w, div));
      const float mm = fmaxf(1e-12f, __fsub_rn(1.0f, fmaxf(fmaxf(c.x, c.y), c.z)));
      c.x = __fdiv_rn(c.x, mm);
      c.y = __fdiv_rn(c.y, mm);
      c.z = __fdiv_rn(c.z, mm);
      outF = c;
      outU = R.format == 2 ? packUnorm4x8(linearToSrgb(c.x), linearToSrgb(c.y), linearToSrgb(c.z), c.w)
                           : packUnorm4x8(c.x, c.y, c.z, c.w);
      have = true;
    }
    if (mine) {
      __stcg(&R.fb.accum[i], acc);
      if (R.format == 0) {
        R.fb.outF32[i] = outF;
        if (R.fb.outMirror)
          __stcs(reinterpret_cast<float4 *>(R.fb.outMirror) + i, outF);
      } else {
        R.fb.outU32[i] = outU;
        if (R.fb.outMirror)
          __stcs(reinterpret_cast<uint32_t *>(R.fb.outMirror) + i, outU);
      }
    }
  }
}

// One step of this warp's share of the compositing.  The tiles of the regions this rank owns form ONE queue in region
// order; a warp draws a ticket from it (one unconditional atomicAdd — no retry, no contention beyond that) and KEEPS
// the ticket until the region it belongs to has been flagged by every rank.  The check never blocks, so during the
// march a warp looks at its pending ticket between two tiles and otherwise keeps marching; regions complete in queue
// order on every rank, so tickets mature in the order they were drawn.  After the tile queue is empty the warp spins
// on what it still holds.  Returns 0 = holding a ticket whose region is not complete yet, 1 = one tile composited,
// 2 = nothing held and the queue is exhausted.
// (History, profiles/r02_sort_last_fused.md: claiming with a CAS after the readiness check made every warp race for the
// same item — 2 us per item, 1 ms per frame at N = 8; per-region counters walked by a per-warp cursor cost every warp
// a dependent load per region — 115 us of tail at N = 2.)
constexpr uint32_t kNoTicket = 0xffffffffu;
__device__ __forceinline__ int slabCompositeStep(const SlabFrameLaunch &S, uint32_t &ticket, uint32_t nItems,
    uint32_t nTiles, int lane)
{
  const PartialLaunch &P = S.m;
  if (ticket == kNoTicket) {
    uint32_t t = kNoTicket;
    if (lane == 0) {
      t = *((volatile unsigned int *)&P.sched[3]);
      if (t < nItems)
        t = atomicAdd(&P.sched[3], 1u);
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nItems)
      return 2;
    ticket = t;
  }
  const uint32_t region = (ticket / S.tilesPerRegion) * S.nRanks + S.rank;
  bool ready = true;
  if ((uint32_t)lane < S.nRanks)
    ready = (int)(*((volatile const unsigned int *)&S.myRegionFlags[(size_t)region * kMaxSlabs + lane]) - S.seq) >= 0;
  if (!__all_sync(0xffffffffu, ready))
    return 0;
  const uint32_t tile = region * S.tilesPerRegion + ticket % S.tilesPerRegion;
  ticket = kNoTicket;
  if (tile < nTiles) {
    const uint32_t tyIdx = P.tileY0 + tile / P.tilesW, txIdx = P.tileX0 + tile % P.tilesW;
    const uint32_t px = txIdx * kTileW + (lane % kTileW), py = tyIdx * kTileH + (lane / kTileW);
    if (px < P.width && py < P.height)
      compositePixel(S.c, px, py, true);
  }
  return 1;
}

template <bool SKIP>
__global__ void __launch_bounds__(kBlockThreads, 2) dvrSlabFrameKernel(const __grid_constant__ SlabFrameLaunch S)
{
  const PartialLaunch &P = S.m;
  __shared__ float4 s_tf[DVR_TF_SIZE];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < DVR_TF_SIZE; i += blockDim.x)
    s_tf[i] = __ldg(&P.inst.v.tf[i]);
  if (threadIdx.x == 0 && S.timing)
    atomicMin(&S.timing[0], globalTimerNs());
  // The partial buffer of this frame was read by the composites of frame seq-2 on every rank: wait for their
  // "resolved" flags (normally long satisfied — our own frame seq-1 already waited for their marches of seq-1).
  if (threadIdx.x == 0 && S.seq > 2u)
    for (uint32_t p = 0; p < S.nRanks; ++p)
      if (!spinUntil(&S.myResolved[p], S.seq - 2u, S.c.sync.errorFlag))
        break;
  __syncthreads();

  MarchStats st{0ull, 0ull};
  const uint32_t nTiles = P.tilesW * P.tilesH;
  const bool centered = P.integrator == DVR_INTEGRATOR_RAYCAST;
  const size_t nBg = S.bgPixelEnd > S.bgPixelBegin ? S.bgPixelEnd - S.bgPixelBegin : 0;
  const uint32_t nChunks = (S.debugFlags & 2u) ? 0u : (uint32_t)((nBg + 255) / 256);
  const uint32_t nOwned = S.nRegions > S.rank ? (S.nRegions - S.rank + S.nRanks - 1u) / S.nRanks : 0u;
  const uint32_t nItems = (S.debugFlags & 4u) ? 0u : nOwned * S.tilesPerRegion;
  uint32_t ticket = kNoTicket; // the composite item this warp holds (slabCompositeStep)
  bool bgLeft = nChunks > 0u, itemsLeft = nItems > 0u;

  // ---- march; between two tiles every warp also takes one background chunk.  Composite items are drawn once the tile
  // queue is empty: the warps that finish early then hold the tickets of the regions still being marched and composite
  // them the moment the last rank's flag arrives.  (DVR_B200_SLAB_DEBUG bit 8 also looks at the held ticket between
  // two tiles — it stretched the march more than it shortened the tail: 1878 vs 1901 frames/s at N = 2, 5092 vs 5412
  // at N = 8, profiles/r02_sort_last_fused.md)
  for (uint32_t tile = nextTile(P.sched, lane); tile < nTiles; tile = nextTile(P.sched, lane)) {
    const uint32_t tyIdx = P.tileY0 + tile / P.tilesW, txIdx = P.tileX0 + tile % P.tilesW;
    const uint32_t px = txIdx * kTileW + (lane % kTileW), py = tyIdx * kTileH + (lane / kTileW);
    if (px < P.width && py < P.height) {
      Philox rng;
      rng.init((unsigned long long)(int)(py * P.width + px), (unsigned long long)P.frameID * 512ull);
      const float4 r = rng.uniform4();
      const float sx = __fmul_rn(centered ? (float)px : __fadd_rn((float)px, r.x), P.invW);
      const float sy = __fmul_rn(centered ? (float)py : __fadd_rn((float)py, r.y), P.invH);
      float3 org, dir;
      cameraCreateRay(P.cam, sx, sy, r.z, r.w, org, dir);
      float3 color = f3(0.f, 0.f, 0.f);
      float opacity = 0.f;
      uint32_t objID = ~0u, instID = ~0u;
      bool anyHit = false;
      const float depth = rayMarchAllVolumes<SKIP, true, false, true, FIELD_STRUCTURED>(&P.inst, 1, TfSelectSingle{s_tf},
          org, dir, FLT_MAX, P.invSamplingRate, rng, color, opacity, objID, instID, st, nullptr, anyHit);
      const uint32_t idx = px + py * P.width;
      P.partialRgba[idx] = make_float4(color.x, color.y, color.z, opacity);
      P.partialDepth[idx] = fminf(1e30f, depth);
    }
    __syncwarp();
    if (lane == 0 && !(S.debugFlags & 1u)) {
      const uint32_t region = tile / S.tilesPerRegion;
      const uint32_t inRegion = min(S.tilesPerRegion, nTiles - region * S.tilesPerRegion);
      // the tile's stores before the count, at device scope.  A RELEASE-only atomic: __threadfence() compiles to
      // MEMBAR.SC.GPU + CCTL.IVALL (an L1 invalidation under the warps that are still marching)
      unsigned int counted;
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;"
                   : "=r"(counted) : "l"(&S.regionDone[region]), "r"(1u) : "memory");
      if (counted == inRegion - 1u) {
        __threadfence_system(); // every tile of the region is in L2; release system-wide before the flag
        *((volatile unsigned int *)&S.regionFlags[region % S.nRanks][(size_t)region * kMaxSlabs + S.rank]) = S.seq;
        if (S.timing)
          atomicMax(&S.timing[6], globalTimerNs());
      }
    }
    __syncwarp();
    if (bgLeft && !(S.debugFlags & 16u)) {
      uint32_t chunk = 0;
      if (lane == 0)
        chunk = atomicAdd(&P.sched[2], 1u);
      chunk = __shfl_sync(0xffffffffu, chunk, 0);
      if (chunk < nChunks)
        slabBackgroundChunk(S, chunk, lane);
      else
        bgLeft = false;
    }
    if (itemsLeft && (S.debugFlags & 8u)) // off by default: measured slower at N = 2, 4 and 8 (profiles/r02_sort_last_fused.md)
      itemsLeft = slabCompositeStep(S, ticket, nItems, nTiles, lane) != 2;
  }
  if (S.timing && lane == 0)
    atomicMax(&S.timing[1], globalTimerNs());

  // ---- the tile queue is empty: what is left of the background strip ...
  while (bgLeft) {
    uint32_t chunk = 0;
    if (lane == 0)
      chunk = atomicAdd(&P.sched[2], 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk < nChunks)
      slabBackgroundChunk(S, chunk, lane);
    else
      bgLeft = false;
  }
  if (S.timing && lane == 0)
    atomicMax(&S.timing[2], globalTimerNs());

  // ---- ... and of the owned regions: the last ones complete when the slowest rank finishes its march
  {
    const long long t0 = clock64();
    // Poll interval: starts at spinSleepNs and doubles up to spinSleepCapNs while the held ticket's region is not
    // complete (every poll of every waiting warp is a strong load of the same few L2 lines, on every SM, next to the
    // warps that are still marching); back to the start after each composited tile.
    unsigned sleepNs = S.spinSleepNs;
    while (itemsLeft) {
      const int r = slabCompositeStep(S, ticket, nItems, nTiles, lane);
      if (r == 2)
        itemsLeft = false;
      else if (r == 1)
        sleepNs = S.spinSleepNs;
      if (r == 0) {
        __nanosleep(sleepNs);
        sleepNs = min(sleepNs * 2u, S.spinSleepCapNs);
        if (clock64() - t0 > 4000000000ll) { // ~2 s: a missing producer must not hang the GPU
          if (lane == 0 && S.c.sync.errorFlag)
            *S.c.sync.errorFlag = 1u;
          break;
        }
      } else if (r == 1 && S.timing && lane == 0)
        atomicMax(&S.timing[5], globalTimerNs());
    }
  }

  // ---- retire
  __syncwarp();
  if (S.timing && lane == 0)
    atomicMax(&S.timing[3], globalTimerNs());
  if (lane == 0) {
    const unsigned int totalWarps = gridDim.x * (blockDim.x >> 5);
    __threadfence_system(); // this warp's composite stores (peer memory) are acknowledged before the count
    const unsigned int done = atomicAdd(&P.sched[1], 1u);
    if (done == totalWarps - 1u) {
      for (uint32_t r = 0; r < S.nRegions; ++r)
        S.regionDone[r] = 0u;
      P.sched[0] = 0u;
      P.sched[1] = 0u;
      P.sched[2] = 0u;
      P.sched[3] = 0u;
      __threadfence_system();
      for (uint32_t p = 0; p < S.nRanks; ++p)
        *((volatile unsigned int *)&S.resolvedFlags[p][S.rank]) = S.seq;
      if (S.waitAllResolved)
        for (uint32_t p = 0; p < S.nRanks; ++p)
          if (!spinUntil(&S.myResolved[p], S.seq, S.c.sync.errorFlag))
            break;
      if (S.timing)
        S.timing[4] = globalTimerNs();
    }
  }
}

// This rank's share of the pixels outside the window as a launch of its own, ahead of the fused frame on the same
// stream — an opt-in (DVR_B200_SLAB_BG_LAUNCH=1).  Inside the fused kernel the chunks sit between march tiles: a
// dependent accumulation read, 8 rows of stores and the index arithmetic per chunk on the critical path of a warp cost
// the march phase 18 us at N = 2 (call Y) and 24 us at N = 8 (call Q).  As a launch of its own the strip takes that
// out of the fused kernel (526 -> 515 us at N = 2) but costs about as much as a serialized launch (525 vs 522 us per
// frame), and with the colour mirrored to pinned host memory its PCIe stores no longer overlap the march (e2e 1603 vs
// 1781 frames/s at N = 2, call Z): the in-kernel placement stays the default.
__global__ void __launch_bounds__(256) dvrSlabBackgroundKernel(const __grid_constant__ SlabFrameLaunch S)
{
  const int lane = threadIdx.x & 31;
  const size_t nBg = S.bgPixelEnd > S.bgPixelBegin ? S.bgPixelEnd - S.bgPixelBegin : 0;
  const uint32_t nChunks = (uint32_t)((nBg + 255) / 256);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t chunk = warp; chunk < nChunks; chunk += nWarps)
    slabBackgroundChunk(S, chunk, lane);
}

static bool slabBackgroundInKernel()
{
  static const bool ownLaunch = [] {
    const char *e = std::getenv("DVR_B200_SLAB_BG_LAUNCH");
    return e && e[0] == '1';
  }();
  return !ownLaunch;
}

int launchSlabFrame(const SlabFrameLaunch &pIn, cudaStream_t s)
{
  SlabFrameLaunch p = pIn;
  if (!(p.debugFlags & 2u) && !slabBackgroundInKernel()) {
    const size_t nBg = p.bgPixelEnd > p.bgPixelBegin ? p.bgPixelEnd - p.bgPixelBegin : 0;
    const unsigned nChunks = (unsigned)((nBg + 255) / 256);
    if (nChunks) {
      const unsigned want = (nChunks + 7u) / 8u, cap = (unsigned)smCount() * 4u;
      dvrSlabBackgroundKernel<<<want < cap ? want : cap, 256, 0, s>>>(p);
      DVR_CUDA(cudaGetLastError());
      countLaunch();
    }
    p.debugFlags |= 2u; // the fused kernel then has no background chunks of its own
  }
  static int bps[2] = {0, 0};
  const int k = p.m.skip ? 1 : 0;
  if (bps[k] == 0) {
    if (k)
      DVR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps[k], dvrSlabFrameKernel<true>, kBlockThreads, 0));
    else
      DVR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps[k], dvrSlabFrameKernel<false>, kBlockThreads, 0));
    if (bps[k] < 1)
      bps[k] = 1;
  }
  // every CTA must be resident: warps of phase 2 spin on flags that other GPUs' resident warps produce, never on
  // work of this grid that has not been scheduled
  const uint32_t grid = (uint32_t)(smCount() * bps[k]);
  if (p.m.skip)
    dvrSlabFrameKernel<true><<<grid, kBlockThreads, 0, s>>>(p);
  else
    dvrSlabFrameKernel<false><<<grid, kBlockThreads, 0, s>>>(p);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

__global__ void dvrScaleVec3Kernel(const float *__restrict__ in, float *__restrict__ out, size_t n, float scale)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[i] * scale;
}

int launchScaleVec3(const float *in, float *out, size_t nPixels, float scale, cudaStream_t s)
{
  const size_t n = nPixels * 3;
  if (n == 0)
    return DVR_OK;
  dvrScaleVec3Kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, n, scale);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

} // namespace dvr
