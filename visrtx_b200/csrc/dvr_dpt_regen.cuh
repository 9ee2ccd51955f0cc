// dvr_dpt_regen.cuh — K2r: the `dpt` renderer as a persistent-lane state machine with pixel regeneration.
//
// The walk of one pixel is the reference's (renderer/DiffusePathTracer_ptx.cu:82-215, gpu/volumeIntegration.h:167-238,
// 352-389, gpu/dda.h:43-121 — same arithmetic and the same Philox draws in the same order as csrc/dvr_dpt.cuh), but the
// schedule is not one warp per 8x4 tile with every lane waiting for the tile's longest path: ncu showed 13.1 of 32 lanes
// active in that kernel (profiles/r02_dpt_ncu.md).  Here every lane owns ONE pixel at a time and runs a flat loop whose
// body is a single step of the tracker — one Woodcock trial in the current majorant cell, or one DDA cell advance — so
// lanes in different cells, bounces and pixels stay converged on the same instructions.  A lane whose path has ended
// parks its result; once DVR_DPT_REFILL lanes of the warp are parked (or nobody is tracking) they accumulate their
// results together and take new pixels from the warp's chunk of the pixel queue (pixels are numbered tile by tile, so
// a warp still works on neighbouring pixels; the next chunk is requested one chunk ahead, the pixel's accumulation
// value is loaded when the pixel is drawn: no service round waits for memory).  A pixel owns its Philox stream (createScreenSample.h:58), so which lane
// traces it, and when, cannot change its result: frames are bit-identical to the tile kernel
// (tests/test_gpu_dpt.py::test_regenerated_lanes_equal_the_tile_kernel).
//
// STATUS: opt-in experiment (DVR_B200_DPT_REGEN=1), not the default.  It is bit-identical to the tile kernel and keeps
// ~28 of 32 lanes busy, but it is slower on the C2 dpt scene at every refill threshold (first version, one atomicAdd per
// service round: refill 1 / 4 / 8 / 16 = 0.773 / 0.694 / 0.637 / 0.555 ms; with the chunk queue and preloaded
// accumulation values: 4 / 8 / 16 / 24 = 0.903 / 0.787 / 0.699 / 0.693 ms; tile kernel 0.488 ms, profiles/r02_dpt_ncu.md):
// the closer the schedule gets to "refill when the whole warp is done", the faster it runs.
//
// Covers what the benchmarked dpt scenes are: one volume instance with an identity transform (structuredRegular or
// NanoVDB).  Several instances, transformed instances and the `test` renderer keep the tile kernel.
#pragma once

#include "dvr_dpt.cuh"
#include "dvr_frame_common.cuh"

namespace dvr {

#ifndef DVR_DPT_REFILL
#define DVR_DPT_REFILL 8 // parked lanes that trigger a service round (accumulate + draw new pixels)
#endif

enum : int
{
  DPT_LANE_EMPTY = 0, // no pixel
  DPT_LANE_DONE = 1,  // path of the current pixel-sample has ended; result not accumulated yet
  DPT_LANE_TRACK = 2, // inside a volume segment
  DPT_LANE_HIT = 3,   // a collision was accepted; scatter / roulette pending
  DPT_LANE_RAY = 4,   // a new ray (primary or scattered) is set up; its segment is not
  DPT_LANE_SAMPLE = 5 // the pixel's next sample is due (primary ray not generated yet)
};

template <int KIND>
struct DptLane
{
  // pixel-sample
  uint32_t px, py;
  int it;
  Philox rng;
  DptPath path;      // persists across the pixel's iterations like PathData in the reference
  float3 primaryDir; // normal channel
  float sx, sy;      // screen coordinate of the sample (background image lookup)
  // current ray (world == object space: identity instance)
  float3 org, dir;
  float tmax; // FLT_MAX for the primary ray, occlusionDistance after a scatter
  // current segment: DDA over the majorant grid (gpu/dda.h:43-121) + Woodcock state of the current cell
  float tLower, rayUpper;
  int cx, cy, cz;
  float3 tnext, dist;
  float t0, t1, t, majorant;
  float3 albedo; // of the accepted collision (DPT_LANE_HIT)
  float4 accum0; // the pixel's accumulation / depth values, loaded when the pixel was drawn (frames after the first)
  float depth0;
  NvdbCache nvCache;
  int state;
};

// sampleDistanceAllVolumes for a single identity instance (volumeIntegration.h:352-389): slab test of the ray interval
// [0, tmax], then the set-up of dda3 + the first cell.  A miss ends the path (Tr stays 1).
template <int KIND>
__device__ __forceinline__ void dptBeginSegment(DptLane<KIND> &L, const VolumeDev &v)
{
  const FieldDev &f = v.f;
  float bt0, bt1;
  if (!intersectVolumeBox(f.boundsLo, f.boundsHi, L.org, L.dir, 0.f, L.tmax, bt0, bt1)) {
    L.state = DPT_LANE_DONE;
    return;
  }
  bt1 = fminf(L.tmax, bt1);
  if (KIND >= FIELD_NANOVDB)
    L.nvCache.reset();
  L.tLower = bt0;
  const float3 oorg = madd3(L.dir, bt0, L.org);
  L.rayUpper = bt1 - bt0;
  const int3 g = v.ddaDims;
  const float3 bl = f.boundsLo, bh = f.boundsHi;
  const float3 ldir = L.dir;
  const float3 rcp = f3(ldir.x != 0.f ? 1.f / ldir.x : 0.f, ldir.y != 0.f ? 1.f / ldir.y : 0.f,
      ldir.z != 0.f ? 1.f / ldir.z : 0.f);
  const float3 lo = (bl - oorg) * rcp, hi = (bh - oorg) * rcp;
  const float3 tnear = f3(fminf(lo.x, hi.x), fminf(lo.y, hi.y), fminf(lo.z, hi.z));
  const float3 tfar = f3(fmaxf(lo.x, hi.x), fmaxf(lo.y, hi.y), fmaxf(lo.z, hi.z));
  const float3 v01 = f3((oorg.x - bl.x) / (bh.x - bl.x), (oorg.y - bl.y) / (bh.y - bl.y), (oorg.z - bl.z) / (bh.z - bl.z));
  L.cx = min(max((int)(v01.x * (float)g.x), 0), g.x - 1);
  L.cy = min(max((int)(v01.y * (float)g.y), 0), g.y - 1);
  L.cz = min(max((int)(v01.z * (float)g.z), 0), g.z - 1);
  L.dist = f3((tfar.x - tnear.x) / (float)g.x, (tfar.y - tnear.y) / (float)g.y, (tfar.z - tnear.z) / (float)g.z);
  L.tnext = f3(__fmaf_rn((float)(ldir.x > 0.f ? L.cx + 1 : g.x - L.cx), L.dist.x, tnear.x),
      __fmaf_rn((float)(ldir.y > 0.f ? L.cy + 1 : g.y - L.cy), L.dist.y, tnear.y),
      __fmaf_rn((float)(ldir.z > 0.f ? L.cz + 1 : g.z - L.cz), L.dist.z, tnear.z));
  L.t0 = 0.f; // fmaxf(rayLower, 0) with rayLower == 0
  L.t1 = fminf(min3(L.tnext), L.rayUpper);
  L.t = L.t0;
  L.majorant = __ldg(&v.ddaMaxOpacities[(size_t)L.cz * g.x * g.y + (size_t)L.cy * g.x + L.cx]);
  L.state = DPT_LANE_TRACK;
}

// makePrimaryRay for the next pixel-sample of the lane's pixel (cameraCreateRay.h:74-81)
template <int KIND>
__device__ __forceinline__ void dptBeginSample(DptLane<KIND> &L, const FrameLaunch &P)
{
  const float4 r = L.rng.uniform4();
  L.sx = __fmul_rn(__fadd_rn((float)L.px, r.x), P.invW);
  L.sy = __fmul_rn(__fadd_rn((float)L.py, r.y), P.invH);
  cameraCreateRay(P.cam, L.sx, L.sy, r.z, r.w, L.org, L.dir);
  L.primaryDir = L.dir;
  L.tmax = FLT_MAX;
  L.state = DPT_LANE_RAY;
}

// A tentative collision was accepted at parameter t of the current segment: the body of dptTracePath's loop after
// sampleDistanceAllVolumes returned (DiffusePathTracer_ptx.cu:150-213 without surfaces)
template <int KIND>
__device__ __forceinline__ void dptCollide(DptLane<KIND> &L, const FrameLaunch &P)
{
  const float3 albedo = L.albedo;
  const float d = L.t + L.tLower;
  if (!(d < L.tmax)) { // sampleDistanceAllVolumes keeps a collision only when it is nearer than the ray's far end
    L.state = DPT_LANE_DONE;
    return;
  }
  if (L.path.depth++ >= P.maxDepth) {
    L.path.Lw = f3(0.f, 0.f, 0.f);
    L.state = DPT_LANE_DONE;
    return;
  }
  const float3 pos = madd3(L.dir, d, L.org);
  L.path.Lw = L.path.Lw * albedo;
  const float Pr = max3(L.path.Lw); // Russian roulette
  if (Pr < .2f) {
    if (L.rng.uniform() > Pr) {
      L.path.Lw = f3(0.f, 0.f, 0.f);
      L.state = DPT_LANE_DONE;
      return;
    }
    L.path.Lw = f3(L.path.Lw.x / Pr, L.path.Lw.y / Pr, L.path.Lw.z / Pr);
  }
  const float3 scatterDir = sampleUnitSphere(L.rng, f3(-L.dir.x, -L.dir.y, -L.dir.z));
  L.org = pos;
  L.dir = scatterDir;
  L.tmax = P.occlusionDistance;
  L.state = DPT_LANE_RAY;
}

// One step of the tracker for a lane in DPT_LANE_TRACK: a Woodcock trial in the current cell (woodcockFunc of
// volumeIntegration.h:188-231) or, when the trial left the cell / the cell is empty, one step of dda3.
template <int KIND>
__device__ __forceinline__ void dptStep(DptLane<KIND> &L, const FrameLaunch &P, const VolumeDev &v,
    const float4 *__restrict__ tf, const float3 halfSpacing, const float invRange)
{
  const FieldDev &f = v.f;
  if (L.majorant > 0.f) {
    L.t = __fmaf_rn(-(logf(1.f - L.rng.uniform()) / L.majorant), f.stepSize, L.t);
    if (!(L.t >= L.t1)) {
      const float3 p = madd3(L.dir, __fadd_rn(L.t, L.tLower), L.org);
      const float s = fieldSample<KIND, false>(f, L.nvCache, fieldCoord<KIND>(f, halfSpacing, p));
      if (!isnan(s)) {
        const float c = __fmul_rn(__fsub_rn(fmaxf(v.vrLower, fminf(s, v.vrUpper)), v.vrLower), invRange);
        const float4 co = tfLookup(tf, c);
        const float u = L.rng.uniform();
        if (co.w >= u * L.majorant) { // accepted: the lane parks until the next service round
          L.albedo = f3(co.x, co.y, co.z);
          L.state = DPT_LANE_HIT;
        }
      }
      return; // next trial in the same cell
    }
  }
  // dda3: step to the next cell; leaving the grid ends the segment without a collision
  const int3 g = v.ddaDims;
  const float t_closest = min3(L.tnext);
  bool left = false;
  if (L.tnext.x == t_closest) {
    L.tnext.x += L.dist.x;
    L.cx += L.dir.x > 0.f ? 1 : -1;
    left = L.cx == (L.dir.x > 0.f ? g.x : -1);
  }
  if (!left && L.tnext.y == t_closest) {
    L.tnext.y += L.dist.y;
    L.cy += L.dir.y > 0.f ? 1 : -1;
    left = L.cy == (L.dir.y > 0.f ? g.y : -1);
  }
  if (!left && L.tnext.z == t_closest) {
    L.tnext.z += L.dist.z;
    L.cz += L.dir.z > 0.f ? 1 : -1;
    left = L.cz == (L.dir.z > 0.f ? g.z : -1);
  }
  if (left) {
    L.state = DPT_LANE_DONE;
    return;
  }
  L.t0 = L.t1;
  L.t1 = fminf(min3(L.tnext), L.rayUpper);
  L.t = L.t0;
  L.majorant = __ldg(&v.ddaMaxOpacities[(size_t)L.cz * g.x * g.y + (size_t)L.cy * g.x + L.cx]);
}

#ifndef DVR_OCC_DPT_REGEN
#define DVR_OCC_DPT_REGEN 3 // structured fields: 80 registers; the tile kernel measured 2 / 3 / 4 CTAs per SM = 0.547 / 0.488 / 0.539 ms
#endif

template <int KIND>
__global__ void __launch_bounds__(kBlockThreads, KIND >= FIELD_NANOVDB ? 2 : DVR_OCC_DPT_REGEN) dvrDptRegenKernel(const __grid_constant__ FrameLaunch P)
{
  __shared__ float4 s_tf[DVR_TF_SIZE];
  const int lane = threadIdx.x & 31;
  const VolumeDev &v = P.inl[0].v;
  for (int i = threadIdx.x; i < DVR_TF_SIZE; i += blockDim.x)
    s_tf[i] = __ldg(&v.tf[i]);
  __syncthreads();

  const float3 halfSpacing = 0.5f * v.f.spacing;
  const float invRange = __fdiv_rn(1.0f, __fsub_rn(v.vrUpper, v.vrLower));
  const bool initFrame = P.frameID == 0 && P.checkerboardID <= 0;
  const AccumCtx actx{P.width, P.height, P.format, P.frameID, P.checkerboardID, P.fb};
  const uint32_t nPix = P.tilesW * P.tilesH * 32u; // pixels of the scheduled tile window, numbered tile by tile
  const unsigned laneLt = (1u << lane) - 1u;

  DptLane<KIND> L;
  L.state = DPT_LANE_EMPTY;
  L.px = L.py = 0u;
  L.it = 0;
  bool more = true; // the pixel queue is not exhausted (warp-uniform)
  constexpr uint32_t kChunk = 32u;
  uint32_t qNext = 0u, qEnd = 0u; // the warp's current chunk of pixel numbers (warp-uniform)
  uint32_t prefetched = 0u;       // lane 0: base of the next chunk
  if (lane == 0)
    prefetched = atomicAdd(&P.sched[0], kChunk);

  while (true) {
    const unsigned tracking = __ballot_sync(0xffffffffu, L.state == DPT_LANE_TRACK);
    const int nTrack = __popc(tracking);
    if (nTrack == 0 || 32 - nTrack >= DVR_DPT_REFILL) {
      // ---- service round (a): accepted collisions -> roulette + scatter (a new ray) or the end of the path
      if (L.state == DPT_LANE_HIT)
        dptCollide(L, P);
      // ---- (b): ended paths -> frame; the pixel's next sample (a new ray), or the lane becomes free
      if (L.state == DPT_LANE_DONE) {
        // DiffusePathTracer_ptx.cu:205-215: colour = Lw * ambient (the background when nothing scattered), alpha 1;
        // depth / ids are never set by the reference's loop; albedo channel = background, normal = primary direction
        const float4 bgd = backgroundAt(P.bgTex, P.background, L.sx, L.sy);
        const float3 c = L.path.depth ? L.path.Lw * P.ambientIntensity : f3(bgd.x, bgd.y, bgd.z);
        const bool pre = !initFrame && L.it == 0; // later samples of the pixel read what this lane just wrote
        accumResults(actx, L.px, L.py, make_float4(c.x, c.y, c.z, 1.f), FLT_MAX, f3(bgd.x, bgd.y, bgd.z), L.primaryDir,
            ~0u, ~0u, ~0u, L.it, initFrame && L.it == 0, pre ? &L.accum0 : nullptr, pre && P.fb.depth ? &L.depth0 : nullptr);
        L.state = ++L.it < P.numIterations ? DPT_LANE_SAMPLE : DPT_LANE_EMPTY;
      }
      // ---- (c): free lanes take the next pixels of the warp's chunk.  Chunks of kChunk consecutive pixels (one 8x4
      // tile) come from the global counter; the NEXT chunk's base is always already requested (lane 0 holds the
      // atomicAdd's result and nobody reads it before the current chunk is used up), so no lane ever waits for it.
      if (more) {
        const unsigned freeMask = __ballot_sync(0xffffffffu, L.state == DPT_LANE_EMPTY);
        if (freeMask) {
          if (qNext >= qEnd) {
            const uint32_t base = __shfl_sync(0xffffffffu, prefetched, 0);
            if (base >= nPix) {
              more = false;
            } else {
              qNext = base;
              qEnd = min(base + kChunk, nPix);
              if (lane == 0)
                prefetched = atomicAdd(&P.sched[0], kChunk);
            }
          }
          if (more) {
            const uint32_t idx = qNext + (uint32_t)__popc(freeMask & laneLt);
            if (L.state == DPT_LANE_EMPTY && idx < qEnd) {
              const uint32_t tile = idx >> 5, pix = idx & 31u;
              const uint32_t tyIdx = P.tileY0 + tile / P.tilesW, txIdx = P.tileX0 + tile % P.tilesW;
              const uint32_t lx = txIdx * kTileW + (pix % kTileW), ly = tyIdx * kTileH + (pix / kTileW);
              bool valid = !(P.tileRanks > 1u && ((tyIdx / P.tileBand) % P.tileRanks) != P.tileRank);
              valid = valid && lx < P.launchW && ly < P.launchH;
              uint32_t px = lx, py = ly;
              if (P.checkerboardID >= 0) { // createScreenSample.h:38-46
                px = lx * 2u + (uint32_t)(P.checkerboardID & 1);
                py = ly * 2u + (uint32_t)((P.checkerboardID >> 1) & 1);
              }
              valid = valid && px < P.width && py < P.height;
              if (valid) {
                L.px = px;
                L.py = py;
                L.it = 0;
                if (!initFrame) { // consumed by accumResults when the pixel's first sample ends
                  const uint32_t pidx = px + py * P.width;
                  L.accum0 = __ldcg(&P.fb.accum[pidx]);
                  if (P.fb.depth)
                    L.depth0 = __ldcg(&P.fb.depth[pidx]);
                }
                L.rng.init((unsigned long long)(int)(py * P.width + px), (unsigned long long)P.frameID * 512ull);
                L.path = DptPath{0, f3(1.f, 1.f, 1.f)};
                L.state = DPT_LANE_SAMPLE;
              }
            }
            qNext = min(qNext + (uint32_t)__popc(freeMask), qEnd);
          }
        }
      }
      // ---- (d): primary rays of the new samples, then every new ray of this round (scattered, next sample, new
      // pixel) enters its segment together; a ray that misses the bounds ends its path and is accumulated next round
      if (L.state == DPT_LANE_SAMPLE)
        dptBeginSample(L, P);
      if (L.state == DPT_LANE_RAY)
        dptBeginSegment(L, v);
      if (!more && __ballot_sync(0xffffffffu, L.state != DPT_LANE_EMPTY) == 0u)
        break;
    }
    if (L.state == DPT_LANE_TRACK)
      dptStep(L, P, v, s_tf, halfSpacing, invRange);
  }
  retireWarp(P.sched, lane);
}

} // namespace dvr
