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
// Mixed scenes.  Unmodified from the reference: sampleLight (gpu/sampleLight.h), computeAO
// (gpu/computeAO.h), surfaceAttenuation / intersectSurface / intersectVolume (gpu/intersectRay.h), rayMarchVolume with a
// null colour (gpu/volumeIntegration.h:300-306), adjustedMaterialOpacity, epsilonFrom, accumulateValue.  Restated:
//   - the surface branch of the raygen programs and shadeSurface / volumeAttenuation (they live in *_ptx.cu files
//     full of OptiX program entry points)          renderer/DirectLight_ptx.cu:64-218,294-418, Raycast_ptx.cu:60-179
//   - the matte material's callable programs        shaders/MatteShader_ptx.cu:39-80 (constant colour / opacity)
//   - RT-core traversal + sphere intersection + any-hit / closest-hit programs: a loop over all primitives, triangle
//     test in double precision                      scene/Intersectors_ptx.cu:74-98, gpu/populateHit.h:196-368,
//                                                   renderer/DirectLight_ptx.cu:228-266
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "gpu/gpu_util.h"
#include "gpu/intersectRay.h"
#include "gpu/createScreenSample.h"
#include "gpu/volumeIntegration.h"
#include "gpu/sampleLight.h"
#include "gpu/computeAO.h"

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
// ---- surfaces of the flattened world (what the surface TLAS, its SBT records and the registry hold) ---------------
struct RefSurfaceRec
{
  GeometryGPUData geom; // the reference's own record: tri.{vertices,indices,vertexNormals,cullBackfaces} / sphere.*
  uint32_t nPrims;
  vec4 color;    // matte "color"
  float opacity; // matte "opacity"
  AlphaMode alphaMode;
  float cutoff;
  uint32_t surfaceId, instanceId;
  float o2w[12], w2o[12]; // row-major 3x4
  int identity;
};
__device__ const RefSurfaceRec *g_surfs;
__device__ int g_nSurfs;

// payload of a volume shadow ray, DirectLight_ptx.cu:54-58
struct RayAttenuation
{
  const Ray *ray{nullptr};
  float attenuation{0.f};
};

// MatteShader_ptx.cu:39-49 for constant parameters; adjustedMaterialOpacity is evalMaterialParameters.h:392-403 (that
// header needs the ANARI-SDK type helpers, so its three lines are restated)
__device__ float refMatteOpacity(const RefSurfaceRec &r)
{
  const float opacityIn = r.color.w * r.opacity;
  if (r.alphaMode == AlphaMode::OPAQUE)
    return 1.f;
  else if (r.alphaMode == AlphaMode::BLEND)
    return opacityIn;
  return opacityIn < r.cutoff ? 0.f : 1.f;
}

struct RefPrimHit
{
  float t;
  float u, v;
  bool front;
  vec3 n;
};

// one primitive in object space; triangles in double precision (the RT cores are watertight, this is the stand-in)
__device__ bool refIntersectPrim(const RefSurfaceRec &r, uint32_t prim, vec3 o, vec3 d, float tmin, float tmax,
    RefPrimHit &h)
{
  if (r.geom.type == GeometryType::SPHERE) { // intersectSphere, Intersectors_ptx.cu:74-98
    const auto &sd = r.geom.sphere;
    const auto primID = sd.indices ? sd.indices[prim] : prim;
    const auto center = sd.centers[primID];
    const auto radius = sd.radii ? sd.radii[primID] : sd.radius;
    const float rd2 = 1.f / dot(d, d);
    const vec3 CO = center - o;
    const float projCO = dot(CO, d) * rd2;
    const vec3 perp = CO - projCO * d;
    const float l2 = glm::dot(perp, perp);
    const float r2 = radius * radius;
    if (l2 > r2)
      return false;
    const float td = glm::sqrt((r2 - l2) * rd2);
    const float t = projCO - td;
    if (!(t > tmin && t < tmax)) // optixReportIntersection accepts hits inside the ray interval only
      return false;
    h.t = t;
    h.u = h.v = 0.f;
    h.front = true;
    h.n = (o + t * d) - center;
    return true;
  }
  const auto &td = r.geom.tri;
  const uvec3 idx = td.indices ? td.indices[prim] : uvec3(0, 1, 2) + prim * 3u;
  const glm::dvec3 v0(td.vertices[idx.x]), v1(td.vertices[idx.y]), v2(td.vertices[idx.z]);
  const glm::dvec3 O(o), D(d);
  const glm::dvec3 n = glm::cross(v1 - v0, v2 - v0);
  const double dn = glm::dot(D, n);
  if (dn == 0.0)
    return false;
  const double t = glm::dot(v0 - O, n) / dn;
  if (!(t > (double)tmin && t < (double)tmax))
    return false;
  const glm::dvec3 P = O + t * D;
  const double nn = glm::dot(n, n);
  // barycentrics from the edge functions (u: weight of v1, v: weight of v2 — optixGetTriangleBarycentrics)
  const double bu = glm::dot(glm::cross(v0 - v2, P - v2), n) / nn;
  const double bv = glm::dot(glm::cross(v1 - v0, P - v0), n) / nn;
  if (bu < 0.0 || bv < 0.0 || bu + bv > 1.0)
    return false;
  h.t = (float)t;
  h.u = (float)bu;
  h.v = (float)bv;
  h.front = dn < 0.0; // counter-clockwise seen from the ray origin
  return true;
}

__device__ void refObjectRay(const RefSurfaceRec &r, float3 org, float3 dir, vec3 &o, vec3 &d)
{
  o = vec3(org.x, org.y, org.z);
  d = vec3(dir.x, dir.y, dir.z);
  if (!r.identity) {
    const float *m = r.w2o;
    const vec3 a = o, b = d;
    o = vec3(m[0] * a.x + m[1] * a.y + m[2] * a.z + m[3], m[4] * a.x + m[5] * a.y + m[6] * a.z + m[7],
        m[8] * a.x + m[9] * a.y + m[10] * a.z + m[11]);
    d = vec3(m[0] * b.x + m[1] * b.y + m[2] * b.z, m[4] * b.x + m[5] * b.y + m[6] * b.z,
        m[8] * b.x + m[9] * b.y + m[10] * b.z);
  }
}

// surface TLAS trace: closest hit (+ populateSurfaceHit / computeTangentSpace) or, for shadow rays, __anyhit__shadow
__device__ void refSurfaceTrace(ScreenSample &ss, float3 org, float3 dir, float tmin, float tmax, void *data,
    unsigned rayFlags)
{
  const bool shadow = (rayFlags & OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT) != 0u;
  const bool anyhitOff = (rayFlags & OPTIX_RAY_FLAG_DISABLE_ANYHIT) != 0u;
  const bool cullBF = (rayFlags & OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES) != 0u;
  int bestS = -1;
  uint32_t bestP = 0;
  RefPrimHit best{};
  for (int si = 0; si < g_nSurfs; ++si) {
    const RefSurfaceRec &r = g_surfs[si];
    vec3 o, d;
    refObjectRay(r, org, dir, o, d);
    for (uint32_t p = 0; p < r.nPrims; ++p) {
      RefPrimHit h;
      if (!refIntersectPrim(r, p, o, d, tmin, tmax, h))
        continue;
      if (shadow) { // DirectLight_ptx.cu:232-250
        float &a = *(float *)data;
        accumulateValue(a, refMatteOpacity(r), a);
        if (a >= 0.99f)
          return;
        continue;
      }
      const bool isTri = r.geom.type == GeometryType::TRIANGLE;
      if (isTri && !h.front && (cullBF || (!anyhitOff && r.geom.tri.cullBackfaces))) // ray flag / cullbackFaces()
        continue;
      tmax = h.t;
      best = h;
      bestS = si;
      bestP = p;
    }
  }
  if (shadow || bestS < 0)
    return;
  const RefSurfaceRec &r = g_surfs[bestS];
  SurfaceHit &hit = *(SurfaceHit *)data;
  const vec3 worg(org.x, org.y, org.z), wdir(dir.x, dir.y, dir.z);
  hit.foundHit = true;
  hit.instance = nullptr;
  hit.geometry = &r.geom;
  hit.material = (const MaterialGPUData *)&r; // opaque handle back to the record
  hit.t = best.t;
  hit.hitpoint = worg + (best.t * wdir);
  hit.uvw = vec3(1.f - best.u - best.v, best.u, best.v);
  hit.primID = bestP;
  hit.objID = r.surfaceId;
  hit.instID = r.instanceId;
  hit.epsilon = epsilonFrom(hit.hitpoint, wdir, best.t);
  // computeTangentSpace, populateHit.h:196-330
  if (r.geom.type == GeometryType::TRIANGLE) {
    const auto &td = r.geom.tri;
    const uvec3 idx = td.indices ? td.indices[bestP] : uvec3(0, 1, 2) + bestP * 3u;
    const vec3 v0 = td.vertices[idx.x], v1 = td.vertices[idx.y], v2 = td.vertices[idx.z];
    hit.Ng = normalize(cross(v1 - v0, v2 - v0));
    if (!best.front)
      hit.Ng = -hit.Ng;
    if (td.vertexNormals != nullptr) {
      const vec3 b = hit.uvw;
      hit.Ns = b.x * td.vertexNormals[idx.x] + b.y * td.vertexNormals[idx.y] + b.z * td.vertexNormals[idx.z];
    } else
      hit.Ns = hit.Ng;
    hit.Ns = normalize(hit.Ns);
    if (dot(hit.Ng, hit.Ns) < 0.f)
      hit.Ns = -hit.Ns;
  } else
    hit.Ng = hit.Ns = best.n;
  if (!r.identity) { // optixTransformNormalFromObjectToWorldSpace: (world -> object)^T
    const float *m = r.w2o;
    const vec3 g = hit.Ng, n = hit.Ns;
    hit.Ng = vec3(m[0] * g.x + m[4] * g.y + m[8] * g.z, m[1] * g.x + m[5] * g.y + m[9] * g.z,
        m[2] * g.x + m[6] * g.y + m[10] * g.z);
    hit.Ns = vec3(m[0] * n.x + m[4] * n.y + m[8] * n.z, m[1] * n.x + m[5] * n.y + m[9] * n.z,
        m[2] * n.x + m[6] * n.y + m[10] * n.z);
  }
  hit.Ng = normalize(hit.Ng);
  hit.Ns = normalize(hit.Ns);
}

__device__ void refgpu_shim_trace(unsigned long long, float3 org, float3 dir, float tmin, float tmax, unsigned ssHi,
    unsigned ssLo, unsigned dataHi, unsigned dataLo, unsigned bvhSelection, unsigned rayFlags)
{
  ScreenSample &ss = *(ScreenSample *)detail::unpackPointer(ssHi, ssLo);
  if (bvhSelection != 0u) {
    refSurfaceTrace(ss, org, dir, tmin, tmax, detail::unpackPointer(dataHi, dataLo), rayFlags);
    return;
  }
  const FrameGPUData &fd = *ss.frameData;
  if (rayFlags & OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT) {
    // volume shadow ray: __anyhit__shadow's volume branch (DirectLight_ptx.cu:251-266) for every volume box the ray
    // enters (instance order; the intersection program's lastVolID test reads the RayAttenuation payload as a
    // VolumeHit there — undefined, taken as "no match")
    RayAttenuation &ra = *(RayAttenuation *)detail::unpackPointer(dataHi, dataLo);
    for (int i = 0; i < (int)fd.world.numVolumeInstances; ++i) {
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
        continue;
      const box1 rayt{tmin, tmax};
      t.lower = clamp(t.lower, rayt);
      t.upper = clamp(t.upper, rayt);
      VolumeHit vh;
      vh.foundHit = true;
      vh.volume = &vd;
      vh.instance = &inst;
      vh.localRay.org = lo;
      vh.localRay.dir = ld;
      vh.localRay.t.lower = t.lower;
      vh.localRay.t.upper = t.upper;
      rayMarchVolume(ss, vh, ra.attenuation, fd.renderer.inverseVolumeSamplingRate);
      if (!(ra.attenuation < 0.99f))
        return; // hit accepted: the ray ends
    }
    return;
  }
  VolumeHit &hit = *(VolumeHit *)detail::unpackPointer(dataHi, dataLo);
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

// ---- mixed scenes: the surface branch of the raygen programs -------------------------------

// volumeAttenuation, DirectLight_ptx.cu:64-71
__device__ float refVolumeAttenuation(ScreenSample &ss, Ray r)
{
  RayAttenuation ra;
  ra.ray = &r;
  intersectVolume(ss, r, 1 /*RayType::SHADOW*/, &ra, OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT);
  return ra.attenuation;
}

// shadeSurface, DirectLight_ptx.cu:73-218, with the matte material's callables (MatteShader_ptx.cu:39-80) in place
// of optixDirectCall.  Matte's nextRay is the zero vector, so the bounce loop leaves before its first trace.
__device__ vec4 refShadeSurface(ScreenSample &ss, const Ray &ray, const SurfaceHit &hit)
{
  const auto &rendererParams = frameData.renderer;
  const auto &directLightParams = rendererParams.params.directLight;
  auto &world = frameData.world;
  const RefSurfaceRec &rec = *(const RefSurfaceRec *)hit.material;

  vec3 shadePoint = hit.hitpoint + (hit.epsilon * hit.Ns);

  const float aoFactor = directLightParams.aoSamples > 0
      ? computeAO(ss, ray, 1 /*RayType::SHADOW*/, hit, rendererParams.occlusionDistance, directLightParams.aoSamples)
      : 1.f;

  vec3 contrib = vec3(0.0f);
  const vec3 baseColor = vec3(rec.color); // __direct_callable__init
  const float opacity = refMatteOpacity(rec);

  if (rendererParams.ambientIntensity > 0.0f)
    contrib = rendererParams.ambientColor * rendererParams.ambientIntensity * baseColor; // evaluateTint

  for (size_t i = 0; i < world.numLightInstances; i++) {
    auto *inst = world.lightInstances + i;
    if (!inst)
      continue;
    for (size_t l = 0; l < inst->numLights; l++) {
      const auto lightSample = sampleLight(ss, hit, inst->indices[l], inst->xfm);
      if (lightSample.pdf == 0.0f)
        continue;
      const Ray shadowRay = {
          shadePoint,
          lightSample.dir,
          {0.0f, lightSample.dist},
      };
      const float surface_o = 1.f - surfaceAttenuation(ss, shadowRay, 1 /*RayType::SHADOW*/);
      const float volume_o = 1.f - refVolumeAttenuation(ss, shadowRay);
      const float attenuation = surface_o * volume_o;
      // __direct_callable__shadeSurface
      float NdotL = fmaxf(0.0f, dot(hit.Ns, lightSample.dir));
      const vec3 thisLightContrib = baseColor * float(M_1_PI) * NdotL * lightSample.radiance / lightSample.pdf;
      if (glm::any(glm::isnan(thisLightContrib)))
        continue;
      contrib += thisLightContrib * attenuation;
    }
  }
  contrib *= aoFactor;
  return vec4(contrib, opacity);
}

// centerPixel != 0: Raycast_ptx.cu:60-179; else DirectLight_ptx.cu:294-418
__global__ void refgpu_raygen_scene(int centerPixel)
{
  auto &rendererParams = frameData.renderer;
  auto ss = createScreenSample(frameData);
  if (pixelOutOfFrame(ss.pixel, frameData.fb))
    return;
  const int iters = centerPixel ? 1 : frameData.renderer.numIterations;
  for (int i = 0; i < iters; i++) {
    auto ray = makePrimaryRay(ss, centerPixel != 0);
    float tmax = ray.t.upper;

    SurfaceHit surfaceHit;
    vec3 outputColor(0.f);
    vec3 outputNormal = ray.dir;
    float outputOpacity = 0.f;
    float depth = 1e30f;
    uint32_t primID = ~0u;
    uint32_t objID = ~0u;
    uint32_t instID = ~0u;
    bool firstHit = true;

    while (outputOpacity < 0.99f) {
      ray.t.upper = tmax;
      surfaceHit.foundHit = false;
      intersectSurface(ss, ray, 0 /*RayType::PRIMARY*/, &surfaceHit, primaryRayOptiXFlags(rendererParams));

      vec3 color(0.f);
      float opacity = 0.f;

      if (surfaceHit.foundHit) {
        uint32_t vObjID = ~0u;
        uint32_t vInstID = ~0u;
        const float vDepth = rayMarchAllVolumes(ss, ray, 0, surfaceHit.t, rendererParams.inverseVolumeSamplingRate,
            color, opacity, vObjID, vInstID);

        if (firstHit) {
          const bool volumeFirst = vDepth < surfaceHit.t;
          if (volumeFirst) {
            outputNormal = -ray.dir;
            depth = vDepth;
            primID = 0;
            objID = vObjID;
            instID = vInstID;
          } else {
            outputNormal = centerPixel ? surfaceHit.Ng : surfaceHit.Ns;
            depth = surfaceHit.t;
            primID = computeGeometryPrimId(surfaceHit);
            objID = surfaceHit.objID;
            instID = surfaceHit.instID;
          }
          firstHit = false;
        }

        if (centerPixel) { // Raycast_ptx.cu:120-135
          const RefSurfaceRec &rec = *(const RefSurfaceRec *)surfaceHit.material;
          auto materialBaseColor = vec3(rec.color);
          auto materialOpacity = refMatteOpacity(rec);
          const auto lighting = glm::abs(glm::dot(ray.dir, surfaceHit.Ns)) * rendererParams.ambientColor;
          accumulateValue(color, materialBaseColor * lighting, opacity);
          accumulateValue(opacity, materialOpacity, opacity);
        } else { // DirectLight_ptx.cu:359-368
          const vec4 shadingResult = refShadeSurface(ss, ray, surfaceHit);
          if (glm::any(glm::isnan(vec3(shadingResult)))) {
            color = vec3(0.f);
            opacity = 0.f;
          } else {
            color = vec3(shadingResult);
            opacity = shadingResult.w;
          }
          accumulateValue(color, vec3(shadingResult), opacity);
          accumulateValue(opacity, shadingResult.w, opacity);
        }

        color *= opacity;
        accumulateValue(outputColor, color, outputOpacity);
        accumulateValue(outputOpacity, opacity, outputOpacity);

        ray.t.lower = surfaceHit.t + surfaceHit.epsilon;
      } else {
        uint32_t vObjID = ~0u;
        uint32_t vInstID = ~0u;
        const float volumeDepth = rayMarchAllVolumes(ss, ray, 0, ray.t.upper,
            rendererParams.inverseVolumeSamplingRate, color, opacity, vObjID, vInstID);

        if (firstHit) {
          depth = min(depth, volumeDepth);
          primID = 0;
          objID = vObjID;
          instID = vInstID;
        }

        color *= opacity;

        const auto bg = getBackground(frameData, ss.screen, ray.dir);
        accumulateValue(color, vec3(bg), opacity);
        accumulateValue(opacity, bg.w, opacity);
        accumulateValue(outputColor, color, outputOpacity);
        accumulateValue(outputOpacity, opacity, outputOpacity);
        break;
      }
    }

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

template <typename T>
static const T *refUpload(std::vector<void *> &allocs, const T *host, size_t n)
{
  if (!host || !n)
    return nullptr;
  void *p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess)
    return nullptr;
  cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
  allocs.push_back(p);
  return (const T *)p;
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
  // mixed scenes
  std::vector<void *> surfaceAllocs;
  RefSurfaceRec *surfs = nullptr;
  int nSurfs = 0;
  LightGPUData *lights = nullptr;
  InstanceLightGPUData *lightInst = nullptr;
  DeviceObjectIndex *lightIdx = nullptr;
  int nLights = 0;
  vec3 ambientColor{1.f};
  float ambientIntensity = 0.f;
  float occlusionDistance = 1e20f;
  int aoSamples = 1;
  bool cullTriangleBF = false;
  bool hasSceneParams = false;
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

// Triangle / Sphere ::gpuData + Surface / Matte gpuData + the instance transform (the flattened world's surfaces)
int refgpu_scene_set_surfaces(RefScene *s, const DvrSurfaceDesc *d, uint32_t n)
{
  for (void *p : s->surfaceAllocs)
    cudaFree(p);
  s->surfaceAllocs.clear();
  s->surfs = nullptr;
  s->nSurfs = 0;
  if (!n)
    return 0;
  std::vector<RefSurfaceRec> recs(n);
  for (uint32_t i = 0; i < n; ++i) {
    RefSurfaceRec &r = recs[i];
    std::memset((void *)&r, 0, sizeof(r));
    const bool tri = d[i].geometryType == DVR_GEOMETRY_TRIANGLE;
    r.nPrims = d[i].index ? d[i].nPrimitives : (tri ? d[i].nVertices / 3u : d[i].nVertices);
    r.geom.primitiveId = refUpload(s->surfaceAllocs, d[i].primitiveId, r.nPrims);
    if (tri) {
      r.geom.type = GeometryType::TRIANGLE;
      r.geom.tri.vertices = (const vec3 *)refUpload(s->surfaceAllocs, d[i].vertexPosition, (size_t)d[i].nVertices * 3);
      r.geom.tri.indices = (const uvec3 *)refUpload(s->surfaceAllocs, d[i].index, (size_t)r.nPrims * 3);
      r.geom.tri.vertexNormals = (const vec3 *)refUpload(s->surfaceAllocs, d[i].vertexNormal, (size_t)d[i].nVertices * 3);
      r.geom.tri.cullBackfaces = d[i].cullBackfaces != 0;
    } else {
      r.geom.type = GeometryType::SPHERE;
      r.geom.sphere.centers = (const vec3 *)refUpload(s->surfaceAllocs, d[i].vertexPosition, (size_t)d[i].nVertices * 3);
      r.geom.sphere.indices = refUpload(s->surfaceAllocs, d[i].index, (size_t)r.nPrims);
      r.geom.sphere.radii = refUpload(s->surfaceAllocs, d[i].vertexRadius, (size_t)d[i].nVertices);
      r.geom.sphere.radius = d[i].radius;
    }
    r.color = vec4(d[i].color[0], d[i].color[1], d[i].color[2], d[i].color[3]);
    r.opacity = d[i].opacity;
    r.alphaMode = d[i].alphaMode == DVR_ALPHA_OPAQUE ? AlphaMode::OPAQUE
                                                     : (d[i].alphaMode == DVR_ALPHA_BLEND ? AlphaMode::BLEND : AlphaMode::MASK);
    r.cutoff = d[i].alphaCutoff;
    r.surfaceId = d[i].surfaceId;
    r.instanceId = d[i].instanceId;
    std::memcpy(r.o2w, d[i].objectToWorld, sizeof(r.o2w));
    { // world -> object: glm::inverse of the affine matrix (what OptiX derives for the instance)
      const float *m = r.o2w;
      const glm::dmat4 M(m[0], m[4], m[8], 0.0, m[1], m[5], m[9], 0.0, m[2], m[6], m[10], 0.0, m[3], m[7], m[11], 1.0);
      const glm::dmat4 I = glm::inverse(M);
      for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 4; ++col)
          r.w2o[4 * row + col] = (float)I[col][row];
    }
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    r.identity = std::memcmp(r.o2w, ident, sizeof(ident)) == 0;
  }
  s->surfs = const_cast<RefSurfaceRec *>(refUpload(s->surfaceAllocs, recs.data(), recs.size()));
  s->nSurfs = (int)n;
  return s->surfs ? 0 : -3;
}

// Directional / Point ::gpuData, one identity light instance (World.cpp light instances), and the renderer members
// the surface branch reads (Renderer.cpp:152-207, DirectLight.cpp:49-64)
int refgpu_scene_set_lighting(RefScene *s, const DvrSceneParams *p)
{
  if (s->lights) cudaFree(s->lights);
  if (s->lightInst) cudaFree(s->lightInst);
  if (s->lightIdx) cudaFree(s->lightIdx);
  s->lights = nullptr;
  s->lightInst = nullptr;
  s->lightIdx = nullptr;
  s->nLights = (int)p->nLights;
  s->ambientColor = vec3(p->ambientColor[0], p->ambientColor[1], p->ambientColor[2]);
  s->ambientIntensity = p->ambientRadiance;
  s->occlusionDistance = p->occlusionDistance > 0.f ? p->occlusionDistance : 1e20f;
  s->aoSamples = std::clamp(p->ambientSamples, 0, 256);
  s->cullTriangleBF = p->cullTriangleBackfaces != 0;
  s->hasSceneParams = true;
  if (!p->nLights)
    return 0;
  std::vector<LightGPUData> l(p->nLights);
  std::vector<DeviceObjectIndex> idx(p->nLights);
  for (uint32_t i = 0; i < p->nLights; ++i) {
    l[i].color = vec3(p->lights[i].color[0], p->lights[i].color[1], p->lights[i].color[2]);
    const vec3 v(p->lights[i].vec[0], p->lights[i].vec[1], p->lights[i].vec[2]);
    if (p->lights[i].type == DVR_LIGHT_POINT) {
      l[i].type = LightType::POINT;
      l[i].point.position = v;
      l[i].point.intensity = p->lights[i].strength;
    } else {
      l[i].type = LightType::DIRECTIONAL;
      l[i].distant.direction = v;
      l[i].distant.irradiance = p->lights[i].strength;
    }
    idx[i] = (DeviceObjectIndex)i;
  }
  RCK(cudaMalloc(&s->lights, sizeof(LightGPUData) * p->nLights));
  RCK(cudaMalloc(&s->lightIdx, sizeof(DeviceObjectIndex) * p->nLights));
  RCK(cudaMalloc(&s->lightInst, sizeof(InstanceLightGPUData)));
  RCK(cudaMemcpy(s->lights, l.data(), sizeof(LightGPUData) * p->nLights, cudaMemcpyHostToDevice));
  RCK(cudaMemcpy(s->lightIdx, idx.data(), sizeof(DeviceObjectIndex) * p->nLights, cudaMemcpyHostToDevice));
  InstanceLightGPUData li;
  li.indices = s->lightIdx;
  li.numLights = p->nLights;
  li.xfm = mat4(1.f);
  RCK(cudaMemcpy(s->lightInst, &li, sizeof(li), cudaMemcpyHostToDevice));
  return 0;
}

int refgpu_scene_destroy(RefScene *s)
{
  if (!s) return 0;
  cudaFree(s->fields); cudaFree(s->volumes); cudaFree(s->instances); cudaFree(s->volIdx); cudaFree(s->xfms);
  cudaFree(s->camera);
  for (void *p : s->surfaceAllocs)
    cudaFree(p);
  cudaFree(s->lights); cudaFree(s->lightInst); cudaFree(s->lightIdx);
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
  if (scene->hasSceneParams) { // mixed scenes: the renderer members of DvrSceneParams
    fd.renderer.ambientColor = scene->ambientColor;
    fd.renderer.ambientIntensity = scene->ambientIntensity;
    fd.renderer.occlusionDistance = scene->occlusionDistance;
    fd.renderer.params.directLight.aoSamples = scene->aoSamples;
    fd.renderer.params.directLight.lightFalloff = 1.f;
  }
  fd.renderer.params.dpt.maxDepth = p->maxDepth <= 0 ? 5 : (p->maxDepth > 256 ? 256 : p->maxDepth);
  fd.renderer.cullTriangleBF = scene->cullTriangleBF;
  fd.renderer.inverseVolumeSamplingRate = p->inverseVolumeSamplingRate;
  fd.renderer.numIterations = p->checkerboardID >= 0 ? 1 : (p->numIterations > 1 ? p->numIterations : 1);
  fd.renderer.maxRayDepth = 5;
  fd.world.volumeInstances = scene->instances;
  fd.world.numVolumeInstances = scene->n;
  fd.world.hdri = -1;
  fd.world.lightInstances = scene->lightInst;
  fd.world.numLightInstances = scene->nLights ? 1 : 0;
  fd.registry.lights = scene->lights;
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
  const RefSurfaceRec *sf = scene->surfs;
  RCK(cudaMemcpyToSymbolAsync(g_surfs, &sf, sizeof(sf), 0, cudaMemcpyHostToDevice, s));
  RCK(cudaMemcpyToSymbolAsync(g_nSurfs, &scene->nSurfs, sizeof(int), 0, cudaMemcpyHostToDevice, s));
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
  } else if (scene->nSurfs > 0)
    refgpu_raygen_scene<<<grid, block, 0, s>>>(p->integrator == DVR_INTEGRATOR_RAYCAST ? 1 : 0);
  else
    refgpu_raygen<<<grid, block, 0, s>>>(p->integrator == DVR_INTEGRATOR_RAYCAST ? 1 : 0);
  RCK(cudaGetLastError());
  return 0;
}

} // extern "C"
