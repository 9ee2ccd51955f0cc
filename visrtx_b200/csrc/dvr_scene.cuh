// dvr_scene.cuh — surfaces, lights and shadow rays around the volume march.
//
// What the reference does with OptiX programs is done here with one BVH over the primitives of the flattened world
// (world-space boxes, object-space primitive tests) walked by the thread that owns the pixel:
//   primary / closest hit          scene/Intersectors_ptx.cu:74-98 (sphere), RT-core triangles, gpu/populateHit.h:196-368
//   surface shadow rays (any hit)  renderer/DirectLight_ptx.cu:228-250, gpu/intersectRay.h:104-111
//   volume shadow rays             renderer/DirectLight_ptx.cu:64-71,251-266, gpu/volumeIntegration.h:300-306
//   ambient occlusion              gpu/computeAO.h:39-60, gpu/gpu_util.h:190-208
//   light sampling                 gpu/sampleLight.h:54-78 (directional, point)
//   matte shading                  shaders/MatteShader_ptx.cu:39-80, gpu/evalMaterialParameters.h:392-403
#pragma once

#include "dvr_march.cuh"

namespace dvr {

struct BvhNode // 32 bytes
{
  float3 lo;
  uint32_t leftOrFirst; // inner: index of the left child (right = left + 1); leaf: first primitive reference
  float3 hi;
  uint32_t count; // 0 = inner node
};

struct ScenePrimRef
{
  uint32_t surface; // index into SceneDev::surfaces
  uint32_t prim;    // primitive index within the surface's geometry
};

struct SceneSurfaceDev
{
  int geometryType;
  int cullBackfaces;
  const float *vertices;       // packed vec3
  const uint32_t *index;       // uvec3 per triangle / uint per sphere, or null
  const float *normals;        // packed vec3 per vertex, or null
  const float *radii;          // per vertex, or null
  const uint32_t *primitiveId; // or null
  float radius;
  float3 baseColor;
  float opacity; // adjustedMaterialOpacity(color.w * opacity, alphaMode, cutoff): constant per surface
  uint32_t surfaceId, instanceId;
  float o2w[12], w2o[12]; // row-major 3x4
  uint32_t identity;
};

struct LightDev
{
  int type;
  float3 color;
  float3 vec;
  float strength;
};

struct SceneDev
{
  const BvhNode *nodes;
  const ScenePrimRef *prims;
  const SceneSurfaceDev *surfaces;
  uint32_t nNodes, nPrims;
  const LightDev *lights;
  int nLights;
  float3 ambientColor;
  float ambientIntensity;
  float occlusionDistance;
  int aoSamples;
  int cullTriangleBF;
};

// SurfaceHit of gpu/gpu_math.h:118-139, the members the matte path reads
struct SurfaceHitDev
{
  bool found;
  float t;
  float3 hitpoint, Ng, Ns;
  float epsilon;
  uint32_t primID, objID, instID;
  const SceneSurfaceDev *surface;
};

__device__ __forceinline__ float3 cross3(float3 a, float3 b)
{
  return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 ld3(const float *p, uint32_t i) { return f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

// transposed 3x3 of a row-major 3x4: optixTransformNormalFromObjectToWorldSpace multiplies by (world->object)^T
__device__ __forceinline__ float3 xfmNormal(const float *w2o, float3 n)
{
  return f3(w2o[0] * n.x + w2o[4] * n.y + w2o[8] * n.z, w2o[1] * n.x + w2o[5] * n.y + w2o[9] * n.z,
      w2o[2] * n.x + w2o[6] * n.y + w2o[10] * n.z);
}

// what one primitive test reports (object space)
struct PrimHit
{
  float t, u, v;
  bool front;
  float3 n; // sphere: h - centre
};

__device__ __forceinline__ bool intersectTriangle(const SceneSurfaceDev &sd, uint32_t prim, const float3 o,
    const float3 d, float tmin, float tmax, PrimHit &h)
{
  uint32_t i0 = 3u * prim, i1 = i0 + 1u, i2 = i0 + 2u;
  if (sd.index) {
    i0 = sd.index[3u * prim];
    i1 = sd.index[3u * prim + 1u];
    i2 = sd.index[3u * prim + 2u];
  }
  const float3 v0 = ld3(sd.vertices, i0);
  const float3 e1 = ld3(sd.vertices, i1) - v0, e2 = ld3(sd.vertices, i2) - v0;
  const float3 p = cross3(d, e2);
  const float det = dot3(e1, p); // == -dot(d, cross(e1, e2)): > 0 for a front-face (counter-clockwise) hit
  if (det == 0.f)
    return false;
  const float inv = 1.f / det;
  const float3 s = o - v0;
  const float u = dot3(s, p) * inv;
  if (!(u >= 0.f && u <= 1.f))
    return false;
  const float3 q = cross3(s, e1);
  const float v = dot3(d, q) * inv;
  if (!(v >= 0.f && u + v <= 1.f))
    return false;
  const float t = dot3(e2, q) * inv;
  if (!(t > tmin && t < tmax))
    return false;
  h.t = t;
  h.u = u;
  h.v = v;
  h.front = det > 0.f;
  return true;
}

// intersectSphere, scene/Intersectors_ptx.cu:74-98: the near root only
__device__ __forceinline__ bool intersectSphere(const SceneSurfaceDev &sd, uint32_t prim, const float3 o, const float3 d,
    float tmin, float tmax, PrimHit &h)
{
  const uint32_t vi = sd.index ? sd.index[prim] : prim;
  const float3 center = ld3(sd.vertices, vi);
  const float radius = sd.radii ? sd.radii[vi] : sd.radius;
  const float rd2 = 1.f / dot3(d, d);
  const float3 CO = center - o;
  const float projCO = dot3(CO, d) * rd2;
  const float3 perp = CO - projCO * d;
  const float l2 = dot3(perp, perp);
  const float r2 = radius * radius;
  if (l2 > r2)
    return false;
  const float td = sqrtf((r2 - l2) * rd2);
  const float t = projCO - td;
  if (!(t > tmin && t < tmax))
    return false;
  const float3 hp = madd3(d, t, o);
  h.t = t;
  h.u = h.v = 0.f;
  h.front = true;
  h.n = hp - center;
  return true;
}

__device__ __forceinline__ bool intersectPrim(const SceneSurfaceDev &sd, uint32_t prim, const float3 org,
    const float3 dir, float tmin, float tmax, PrimHit &h)
{
  float3 o = org, d = dir;
  if (!sd.identity) {
    o = xfmPoint(sd.w2o, org);
    d = xfmVector(sd.w2o, dir);
  }
  if (sd.geometryType == DVR_GEOMETRY_SPHERE)
    return intersectSphere(sd, prim, o, d, tmin, tmax, h);
  return intersectTriangle(sd, prim, o, d, tmin, tmax, h);
}

__device__ __forceinline__ bool rayBox(const BvhNode &n, const float3 org, const float3 inv, float tmin, float tmax)
{
  const float3 a = (n.lo - org) * inv, b = (n.hi - org) * inv;
  const float tn = fmaxf(fmaxf(fminf(a.x, b.x), fminf(a.y, b.y)), fmaxf(fminf(a.z, b.z), tmin));
  const float tf = fminf(fminf(fmaxf(a.x, b.x), fmaxf(a.y, b.y)), fminf(fmaxf(a.z, b.z), tmax));
  return tn <= tf;
}

constexpr int kBvhStack = 48;

// closest hit with the primary-ray culling rules: OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES when the renderer asks
// for it (gpu/intersectRay.h:80-84), else __anyhit__primary's per-geometry cullBackfaces (gpu/populateHit.h:331-343)
__device__ __forceinline__ void intersectSurfaceClosest(const SceneDev &sc, const float3 org, const float3 dir,
    float tmin, float tmax, SurfaceHitDev &hit)
{
  hit.found = false;
  if (sc.nNodes == 0u)
    return;
  const float3 inv = f3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
  uint32_t stack[kBvhStack];
  int sp = 0;
  stack[sp++] = 0u;
  PrimHit best{};
  uint32_t bestRef = ~0u;
  while (sp > 0) {
    const BvhNode n = sc.nodes[stack[--sp]];
    if (!rayBox(n, org, inv, tmin, tmax))
      continue;
    if (n.count == 0u) {
      if (sp + 2 <= kBvhStack) {
        stack[sp++] = n.leftOrFirst;
        stack[sp++] = n.leftOrFirst + 1u;
      }
      continue;
    }
    for (uint32_t k = 0; k < n.count; ++k) {
      const ScenePrimRef r = sc.prims[n.leftOrFirst + k];
      const SceneSurfaceDev &sd = sc.surfaces[r.surface];
      PrimHit h;
      if (!intersectPrim(sd, r.prim, org, dir, tmin, tmax, h))
        continue;
      if (sd.geometryType == DVR_GEOMETRY_TRIANGLE && !h.front && (sc.cullTriangleBF || sd.cullBackfaces))
        continue;
      tmax = h.t;
      best = h;
      bestRef = n.leftOrFirst + k;
    }
  }
  if (bestRef == ~0u)
    return;
  // populateSurfaceHit + computeTangentSpace, gpu/populateHit.h:196-368
  const ScenePrimRef r = sc.prims[bestRef];
  const SceneSurfaceDev &sd = sc.surfaces[r.surface];
  hit.found = true;
  hit.surface = &sd;
  hit.t = best.t;
  hit.hitpoint = madd3(dir, best.t, org);
  hit.primID = sd.primitiveId ? sd.primitiveId[r.prim] : r.prim;
  hit.objID = sd.surfaceId;
  hit.instID = sd.instanceId;
  { // epsilonFrom, gpu/gpu_util.h:245-250
    const float dm = fmaxf(fmaxf(fabsf(dir.x), fabsf(dir.y)), fabsf(dir.z)) * best.t;
    hit.epsilon = fmaxf(fmaxf(fabsf(hit.hitpoint.x), fabsf(hit.hitpoint.y)), fmaxf(fabsf(hit.hitpoint.z), dm))
        * 0x1.fp-21f;
  }
  float3 Ng, Ns;
  if (sd.geometryType == DVR_GEOMETRY_SPHERE) {
    Ng = Ns = best.n;
  } else {
    uint32_t i0 = 3u * r.prim, i1 = i0 + 1u, i2 = i0 + 2u;
    if (sd.index) {
      i0 = sd.index[3u * r.prim];
      i1 = sd.index[3u * r.prim + 1u];
      i2 = sd.index[3u * r.prim + 2u];
    }
    const float3 v0 = ld3(sd.vertices, i0);
    Ng = normalize3(cross3(ld3(sd.vertices, i1) - v0, ld3(sd.vertices, i2) - v0));
    if (!best.front)
      Ng = f3(-Ng.x, -Ng.y, -Ng.z);
    if (sd.normals) {
      const float b0 = 1.f - best.u - best.v;
      const float3 n0 = ld3(sd.normals, i0), n1 = ld3(sd.normals, i1), n2 = ld3(sd.normals, i2);
      Ns = f3(b0 * n0.x + best.u * n1.x + best.v * n2.x, b0 * n0.y + best.u * n1.y + best.v * n2.y,
          b0 * n0.z + best.u * n1.z + best.v * n2.z);
    } else
      Ns = Ng;
    Ns = normalize3(Ns);
    if (dot3(Ng, Ns) < 0.f)
      Ns = f3(-Ns.x, -Ns.y, -Ns.z);
  }
  if (!sd.identity) {
    Ng = xfmNormal(sd.w2o, Ng);
    Ns = xfmNormal(sd.w2o, Ns);
  }
  hit.Ng = normalize3(Ng);
  hit.Ns = normalize3(Ns);
}

// surfaceAttenuation, gpu/intersectRay.h:104-111 with __anyhit__shadow (DirectLight_ptx.cu:228-250): every primitive
// the ray crosses adds its material opacity (accumulateValue), terminating at 0.99.  No culling on shadow rays.
__device__ __forceinline__ float surfaceAttenuation(const SceneDev &sc, const float3 org, const float3 dir, float tmin,
    float tmax)
{
  float o = 0.f;
  if (sc.nNodes == 0u)
    return o;
  const float3 inv = f3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
  uint32_t stack[kBvhStack];
  int sp = 0;
  stack[sp++] = 0u;
  while (sp > 0) {
    const BvhNode n = sc.nodes[stack[--sp]];
    if (!rayBox(n, org, inv, tmin, tmax))
      continue;
    if (n.count == 0u) {
      if (sp + 2 <= kBvhStack) {
        stack[sp++] = n.leftOrFirst;
        stack[sp++] = n.leftOrFirst + 1u;
      }
      continue;
    }
    for (uint32_t k = 0; k < n.count; ++k) {
      const ScenePrimRef r = sc.prims[n.leftOrFirst + k];
      const SceneSurfaceDev &sd = sc.surfaces[r.surface];
      PrimHit h;
      if (!intersectPrim(sd, r.prim, org, dir, tmin, tmax, h))
        continue;
      o += sd.opacity * (1.f - o);
      if (o >= 0.99f)
        return o;
    }
  }
  return o;
}

// One volume of a shadow ray: what __anyhit__shadow runs per volume box the ray enters — the slab test of
// Intersectors_ptx.cu:248-274 on [0, dist], then rayMarchVolume with a null colour (opacity only; both jitters are
// drawn from the pixel's Philox stream exactly like a primary segment).
template <bool SKIP, typename TfSelect>
__device__ __forceinline__ float volumeAttenuation(const InstanceDev *__restrict__ inst, const int nInst, TfSelect tfOf,
    const float3 org, const float3 dir, const float dist, const float invSamplingRate, Philox &rng)
{
  float attenuation = 0.f;
  MarchStats st{0ull, 0ull};
  for (int i = 0; i < nInst; ++i) {
    const InstanceDev &in = inst[i];
    float3 lo = org, ld = dir;
    if (!in.identity) {
      lo = xfmPoint(in.xfm, org);
      ld = xfmVector(in.xfm, dir);
    }
    float t0, t1;
    if (!intersectVolumeBox(in.v.f.boundsLo, in.v.f.boundsHi, lo, ld, 0.f, dist, t0, t1))
      continue;
    const float tStart = __fmaf_rn(in.v.f.stepSize, rng.uniform(), t0);
    float3 unused = f3(0.f, 0.f, 0.f);
    if (in.v.f.kind == FIELD_NANOVDB_QUANT)
      marchSegment<SKIP, false, false, FIELD_NANOVDB_QUANT>(
          in.v, tfOf(i), lo, ld, tStart, t1, invSamplingRate, rng, unused, attenuation, st, nullptr);
    else if (in.v.f.kind == FIELD_NANOVDB)
      marchSegment<SKIP, false, false, FIELD_NANOVDB>(
          in.v, tfOf(i), lo, ld, tStart, t1, invSamplingRate, rng, unused, attenuation, st, nullptr);
    else
      marchSegment<SKIP, false, false, FIELD_STRUCTURED>(
          in.v, tfOf(i), lo, ld, tStart, t1, invSamplingRate, rng, unused, attenuation, st, nullptr);
    if (attenuation >= 0.99f) // the any-hit program accepts the hit: the shadow ray ends here
      break;
  }
  return attenuation;
}

// shadeSurface of the directLight renderer (DirectLight_ptx.cu:73-218) for a matte material: ambient term, every
// light with its surface- and volume-attenuated shadow ray, ambient occlusion on the sum.  Matte's nextRay is the
// zero vector (MatteShader_ptx.cu:51-57), so the bounce loop ends before its first trace.
template <bool SKIP, typename TfSelect>
__device__ __forceinline__ float4 shadeSurfaceDirectLight(const SceneDev &sc, const InstanceDev *__restrict__ inst,
    const int nInst, TfSelect tfOf, const float3 rayDir, const SurfaceHitDev &hit, const float invSamplingRate,
    Philox &rng)
{
  const float3 shadePoint = madd3(hit.Ns, hit.epsilon, hit.hitpoint);
  float aoFactor = 1.f;
  if (sc.aoSamples > 0) { // computeAO, gpu/computeAO.h:39-60
    float weights = 0.f, hits = 0.f;
    const float3 aoOrg = madd3(hit.Ng, hit.epsilon, hit.hitpoint);
    for (int i = 0; i < sc.aoSamples; ++i) {
      const float4 r = rng.uniform4(); // randomDir, gpu_util.h:190-208
      float3 d = normalize3(f3(2.f * r.x - 1.f, 2.f * r.y - 1.f, 2.f * r.z - 1.f));
      if (!(dot3(d, hit.Ns) > 0.f))
        d = f3(-d.x, -d.y, -d.z);
      const float weight = fmaxf(0.f, dot3(d, hit.Ns));
      weights += weight;
      if (weight != 0.f)
        hits += weight * surfaceAttenuation(sc, aoOrg, d, 0.f, sc.occlusionDistance);
    }
    aoFactor = weights > 0.f ? 1.f - hits / weights : 0.f;
  }
  const SceneSurfaceDev &sd = *hit.surface;
  float3 contrib = f3(0.f, 0.f, 0.f);
  if (sc.ambientIntensity > 0.f)
    contrib = f3(sc.ambientColor.x * sc.ambientIntensity * sd.baseColor.x,
        sc.ambientColor.y * sc.ambientIntensity * sd.baseColor.y,
        sc.ambientColor.z * sc.ambientIntensity * sd.baseColor.z);
  for (int l = 0; l < sc.nLights; ++l) {
    const LightDev &ld = sc.lights[l];
    float3 ldir, radiance;
    float ldist;
    if (ld.type == DVR_LIGHT_POINT) { // samplePointLight
      ldir = ld.vec - hit.hitpoint;
      ldist = sqrtf(dot3(ldir, ldir));
      ldir = normalize3(ldir);
    } else { // sampleDirectionalLight: towards the light
      ldir = f3(-ld.vec.x, -ld.vec.y, -ld.vec.z);
      ldist = __int_as_float(0x7f800000);
    }
    radiance = f3(ld.color.x * ld.strength, ld.color.y * ld.strength, ld.color.z * ld.strength);
    const float surface_o = 1.f - surfaceAttenuation(sc, shadePoint, ldir, 0.f, ldist);
    const float volume_o = 1.f - volumeAttenuation<SKIP>(inst, nInst, tfOf, shadePoint, ldir, ldist, invSamplingRate, rng);
    const float attenuation = surface_o * volume_o;
    const float NdotL = fmaxf(0.f, dot3(hit.Ns, ldir)); // MatteShader_ptx.cu:78-79, pdf == 1
    const float k = 0.318309886183790671538f * NdotL;
    const float3 c = f3(sd.baseColor.x * k * radiance.x, sd.baseColor.y * k * radiance.y, sd.baseColor.z * k * radiance.z);
    if (isnan(c.x) || isnan(c.y) || isnan(c.z))
      continue;
    contrib = f3(contrib.x + c.x * attenuation, contrib.y + c.y * attenuation, contrib.z + c.z * attenuation);
  }
  (void)rayDir;
  return make_float4(contrib.x * aoFactor, contrib.y * aoFactor, contrib.z * aoFactor, sd.opacity);
}

} // namespace dvr
