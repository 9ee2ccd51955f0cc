// dvr_frame_common.cuh — device pieces shared by the frame kernels (dvr_kernels.cu) and the scene kernel
// (dvr_scene.cu): warp-granular tile scheduler, accumResults tail, missed-pixel shading, TF table selectors.
#pragma once

#include "dvr_internal.h"
#include "dvr_march.cuh"

namespace dvr {

// ----------------------------------------------------------------------------------------------
// warp-granular dynamic tile scheduler.  sched[0] = next tile, sched[1] = warps finished.  The last
// warp to leave re-arms both counters, so no memset is needed between frames.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nextTile(unsigned int *sched, int lane)
{
  uint32_t t = 0;
  if (lane == 0)
    t = atomicAdd(&sched[0], 1u);
  return __shfl_sync(0xffffffffu, t, 0);
}

__device__ __forceinline__ void retireWarp(unsigned int *sched, int lane)
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
    }
  }
}

__device__ __forceinline__ unsigned long long warpSum(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// accumResults, gpu/gpu_util.h:393-443, for one pixel-sample.  `init` replaces the cleared
// buffers of Frame::newFrame (Frame.cu:609-647): 0 + x and min(FLT_MAX, x) are written directly.
struct AccumCtx
{
  uint32_t width, height;
  int format, frameID, checkerboardID;
  BuffersDev fb;
};

__device__ __forceinline__ void accumResults(const AccumCtx &P, uint32_t px, uint32_t py, float4 color,
    float depth, float3 albedo, float3 normal, uint32_t primID, uint32_t objID, uint32_t instID,
    int frameIDOffset, bool init, const float4 *preAccum = nullptr, const float *preDepth = nullptr)
{
  // preAccum / preDepth: the pixel's accumulation / depth values loaded earlier by the caller (K2r issues the loads
  // when it draws the pixel, so their latency is covered by the walk), instead of two dependent loads here
  const BuffersDev &fb = P.fb;
  const uint32_t idx = px + py * P.width;
  const int frameID = P.frameID + frameIDOffset;

  // tonemap: v / (1 + max(0, compMax(v)))
  const float m = __fadd_rn(1.0f, fmaxf(0.0f, fmaxf(fmaxf(color.x, color.y), color.z)));
  const float4 tm = make_float4(__fdiv_rn(color.x, m), __fdiv_rn(color.y, m), __fdiv_rn(color.z, m), color.w);

  float4 acc;
  if (init) {
    acc = tm;
  } else {
    acc = preAccum ? *preAccum : __ldcg(&fb.accum[idx]); // streamed once per frame: keep it out of L1, where the field's texels live
    acc.x = __fadd_rn(acc.x, tm.x);
    acc.y = __fadd_rn(acc.y, tm.y);
    acc.z = __fadd_rn(acc.z, tm.z);
    acc.w = __fadd_rn(acc.w, tm.w);
  }
  __stcg(&fb.accum[idx], acc);

  if (fb.albedo) {
    float *a = fb.albedo + 3 * (size_t)idx;
    if (init) {
      a[0] = albedo.x; a[1] = albedo.y; a[2] = albedo.z;
    } else {
      a[0] += albedo.x; a[1] += albedo.y; a[2] += albedo.z;
    }
  }
  if (fb.normal) {
    float *n = fb.normal + 3 * (size_t)idx;
    if (init) {
      n[0] = normal.x; n[1] = normal.y; n[2] = normal.z;
    } else {
      n[0] += normal.x; n[1] += normal.y; n[2] += normal.z;
    }
  }

  bool closer = true;
  if (fb.depth) {
    const float prev = init ? FLT_MAX : (preDepth ? *preDepth : __ldcg(&fb.depth[idx]));
    closer = depth < prev;
    if (closer)
      fb.depth[idx] = depth;
    else if (init)
      fb.depth[idx] = prev;
    if (fb.depthMirror && (closer || init))
      fb.depthMirror[idx] = closer ? depth : prev;
  }
  if (closer) {
    if (fb.primId) fb.primId[idx] = primID;
    if (fb.objId) fb.objId[idx] = objID;
    if (fb.instId) fb.instId[idx] = instID;
  } else if (init) {
    if (fb.primId) fb.primId[idx] = 0u;
    if (fb.objId) fb.objId[idx] = 0u;
    if (fb.instId) fb.instId[idx] = 0u;
  }

  writeOutputColor(fb, P.format, acc, idx, frameID);

  // first checkerboard pass: replicate the colour into the three not-yet-rendered neighbours
  // (gpu_util.h:424-442) and initialise their accumulation state for the passes that follow
  if (P.checkerboardID == 0 && frameID == 0) {
#pragma unroll
    for (int n = 1; n < 4; ++n) {
      const uint32_t ax = px + (n & 1), ay = py + (n >> 1);
      if (ax >= P.width || ay >= P.height)
        continue;
      const uint32_t aidx = ax + ay * P.width;
      writeOutputColor(fb, P.format, acc, aidx, frameID);
      if (init) {
        fb.accum[aidx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fb.depth) fb.depth[aidx] = FLT_MAX;
        if (fb.depth && fb.depthMirror) fb.depthMirror[aidx] = FLT_MAX;
        if (fb.primId) fb.primId[aidx] = 0u;
        if (fb.objId) fb.objId[aidx] = 0u;
        if (fb.instId) fb.instId[aidx] = 0u;
        if (fb.albedo) { float *a = fb.albedo + 3 * (size_t)aidx; a[0] = a[1] = a[2] = 0.f; }
        if (fb.normal) { float *a = fb.normal + 3 * (size_t)aidx; a[0] = a[1] = a[2] = 0.f; }
      }
    }
  }
}

// A pixel none of whose rays can enter a volume: what the march + Raycast_ptx.cu:139-166 produce for a miss, bit for
// bit (colour 0*0 + bg*(1-0) = bg, depth min(1e30, tmax), ids ~0u).  With a constant background no Philox or camera
// work is needed; with a background image the screen coordinate of every pixel-sample is regenerated (each missed
// iteration consumes exactly the four draws of makePrimaryRay, cameraCreateRay.h:74-81).
__device__ __forceinline__ void shadeMissedPixel(const AccumCtx &actx, uint32_t px, uint32_t py, const float4 bgConst,
    cudaTextureObject_t bgTex, bool centered, float invW, float invH, int numIterations, bool initFrame)
{
  Philox rng;
  if (bgTex && !centered)
    rng.init((unsigned long long)(int)(py * actx.width + px), (unsigned long long)actx.frameID * 512ull);
  for (int it = 0; it < numIterations; ++it) {
    float4 bg = bgConst;
    if (bgTex) {
      float sx = (float)px, sy = (float)py;
      if (!centered) {
        const float4 r = rng.uniform4();
        sx = __fadd_rn(sx, r.x);
        sy = __fadd_rn(sy, r.y);
      }
      bg = tex2D<float4>(bgTex, __fmul_rn(sx, invW), __fmul_rn(sy, invH));
    }
    accumResults(actx, px, py, bg, 1e30f, f3(bg.x, bg.y, bg.z), f3(0.f, 0.f, 0.f), 0u, ~0u, ~0u, it,
        initFrame && it == 0);
  }
}

struct TfSelectShared
{
  const float4 *smem;
  const InstanceDev *inst;
  __device__ __forceinline__ const float4 *operator()(int i) const
  {
    return i < kMaxInlineInstances ? smem + i * DVR_TF_SIZE : inst[i].v.tf;
  }
};
struct TfSelectSingle
{
  const float4 *smem;
  __device__ __forceinline__ const float4 *operator()(int) const { return smem; }
};

} // namespace dvr
