// dvr_dpt.cuh — K2: delta (Woodcock) tracking through the majorant grid for the `dpt` renderer.
//
// Reproduces renderer/DiffusePathTracer_ptx.cu:82-215 for a volume-only world, with
// gpu/volumeIntegration.h:167-238,352-389 (_sampleDistance / sampleDistanceAllVolumes),
// gpu/dda.h:43-121 (dda3), gpu/uniformGrid.h:39-50 (projectOnGrid / linearIndex) and
// gpu/gpu_util.h:205-243 (computeOrthonormalBasis / sampleUnitSphere).  The RNG draw order is the
// reference's: per Woodcock step one uniform for the free path and, for a non-NaN sample, one for
// the acceptance test; one for Russian roulette when max(Lw) < 0.2; two for the scatter direction.
//
// The grid is the reference's geometry (ceil(dims/16) cells dividing the field bounds evenly) but its
// content is built correctly: conservative value ranges per cell and majorants from the volume's own
// value range (SURVEY quirks Q7/Q8 are bugs of the reference's build, not of the tracker).
#pragma once

#include "dvr_march.cuh"

namespace dvr {

// gpu_util.h:205-219
__device__ __forceinline__ void orthonormalBasis(const float3 n, float3 &u, float3 &v)
{
  const float sign = n.z >= 0.0f ? 1.0f : -1.0f;
  const float a = -1.0f / (sign + n.z);
  const float b = n.x * n.y * a;
  u = f3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
  v = f3(b, sign + n.y * n.y * a, -n.y);
}

// gpu_util.h:234-243
__device__ __forceinline__ float3 sampleUnitSphere(Philox &rng, const float3 normal)
{
  const float cost = 1.f - 2.f * rng.uniform();
  const float sint = sqrtf(fmaxf(0.f, 1.f - cost * cost));
  const float phi = 2.f * 3.14159265358979323846f * rng.uniform();
  float3 u, v;
  orthonormalBasis(normal, u, v);
  const float sx = sint * cosf(phi), sy = sint * sinf(phi), sz = -cost;
  return f3(u.x * sx + v.x * sy + normal.x * sz, u.y * sx + v.y * sy + normal.y * sz,
      u.z * sx + v.z * sy + normal.z * sz);
}

// _sampleDistance (volumeIntegration.h:167-238) for one volume segment: Woodcock tracking cell by cell
// along the 3-D DDA of gpu/dda.h.  Returns the collision distance along the (world-unit) ray, or the
// segment end; tr = 0 on a collision, 1 otherwise.
template <int KIND>
__device__ __forceinline__ float sampleDistanceSegment(const VolumeDev &v, const float4 *__restrict__ tf,
    const float3 lorg, const float3 ldir, const float tLower, const float tUpper, Philox &rng, float3 &albedo,
    float &extinction, float &tr)
{
  const FieldDev &f = v.f;
  const float stepSize = f.stepSize;
  float t_out = tUpper;
  tr = 1.f;
  const float3 halfSpacing = 0.5f * f.spacing;
  const float vrLo = v.vrLower, vrHi = v.vrUpper;
  const float invRange = __fdiv_rn(1.0f, __fsub_rn(vrHi, vrLo));
  NvdbCache nvCache;
  if (KIND >= FIELD_NANOVDB)
    nvCache.reset();

  // objRay: origin moved to the entry point, interval [0, tUpper - tLower]
  const float3 oorg = madd3(ldir, tLower, lorg);
  const float rayLower = 0.f, rayUpper = tUpper - tLower;

  // dda3(objRay, grid.dims, grid.worldBounds, woodcock)
  const int3 g = v.ddaDims;
  const float3 bl = f.boundsLo, bh = f.boundsHi;
  const float3 rcp = f3(ldir.x != 0.f ? 1.f / ldir.x : 0.f, ldir.y != 0.f ? 1.f / ldir.y : 0.f,
      ldir.z != 0.f ? 1.f / ldir.z : 0.f);
  const float3 lo = (bl - oorg) * rcp, hi = (bh - oorg) * rcp;
  const float3 tnear = f3(fminf(lo.x, hi.x), fminf(lo.y, hi.y), fminf(lo.z, hi.z));
  const float3 tfar = f3(fmaxf(lo.x, hi.x), fmaxf(lo.y, hi.y), fmaxf(lo.z, hi.z));
  // projectOnGrid(ray.org, dims, bounds)
  const float3 v01 = f3((oorg.x - bl.x) / (bh.x - bl.x), (oorg.y - bl.y) / (bh.y - bl.y), (oorg.z - bl.z) / (bh.z - bl.z));
  int cx = min(max((int)(v01.x * (float)g.x), 0), g.x - 1);
  int cy = min(max((int)(v01.y * (float)g.y), 0), g.y - 1);
  int cz = min(max((int)(v01.z * (float)g.z), 0), g.z - 1);
  const float3 dist = f3((tfar.x - tnear.x) / (float)g.x, (tfar.y - tnear.y) / (float)g.y, (tfar.z - tnear.z) / (float)g.z);
  const int sx = ldir.x > 0.f ? 1 : -1, sy = ldir.y > 0.f ? 1 : -1, sz = ldir.z > 0.f ? 1 : -1;
  const int ex = ldir.x > 0.f ? g.x : -1, ey = ldir.y > 0.f ? g.y : -1, ez = ldir.z > 0.f ? g.z : -1;
  float3 tnext = f3(__fmaf_rn((float)(ldir.x > 0.f ? cx + 1 : g.x - cx), dist.x, tnear.x),
      __fmaf_rn((float)(ldir.y > 0.f ? cy + 1 : g.y - cy), dist.y, tnear.y),
      __fmaf_rn((float)(ldir.z > 0.f ? cz + 1 : g.z - cz), dist.z, tnear.z));
  float t0 = fmaxf(rayLower, 0.f);

  while (true) {
    const float t1 = fminf(min3(tnext), rayUpper);
    // ---- woodcockFunc(leafID, t0, t1)
    {
      const float majorant = __ldg(&v.ddaMaxOpacities[(size_t)cz * g.x * g.y + (size_t)cy * g.x + cx]);
      float t = t0;
      while (majorant > 0.f) {
        t = __fmaf_rn(-(logf(1.f - rng.uniform()) / majorant), stepSize, t);
        if (t >= t1)
          break;
        const float3 p = madd3(ldir, __fadd_rn(t, tLower), lorg);
        const float s = fieldSample<KIND, false>(f, nvCache, fieldCoord<KIND>(f, halfSpacing, p));
        if (!isnan(s)) {
          const float c = __fmul_rn(__fsub_rn(fmaxf(vrLo, fminf(s, vrHi)), vrLo), invRange);
          const float4 co = tfLookup(tf, c);
          albedo = f3(co.x, co.y, co.z);
          extinction = co.w;
          const float u = rng.uniform();
          if (extinction >= u * majorant) {
            tr = 0.f;
            t_out = t;
            return t_out + tLower; // stop traversal
          }
        }
      }
    }
    const float t_closest = min3(tnext);
    if (tnext.x == t_closest) {
      tnext.x += dist.x;
      cx += sx;
      if (cx == ex)
        break;
    }
    if (tnext.y == t_closest) {
      tnext.y += dist.y;
      cy += sy;
      if (cy == ey)
        break;
    }
    if (tnext.z == t_closest) {
      tnext.z += dist.z;
      cz += sz;
      if (cz == ez)
        break;
    }
    t0 = t1;
  }
  return t_out + tLower;
}

// sampleDistanceAllVolumes, volumeIntegration.h:352-389: every volume along the ray is tracked; the nearest
// collision wins.  (No lastVolID exclusion needed beyond the reference's: the ray restarts behind each box.)
template <bool SINGLE, int KIND, typename TfSelect>
__device__ __forceinline__ float sampleDistanceAllVolumes(const InstanceDev *__restrict__ inst, const int nInst,
    TfSelect tfOf, const float3 org, const float3 dir, const float tmin, const float tfar, Philox &rng, float3 &albedo,
    float &extinction, float &transmittance)
{
  float rayLower = tmin;
  const float rayUpper = tfar;
  float depth = tfar;
  transmittance = 1.f;
  int last = -1;
  while (true) {
    int best = -1;
    float bt0 = 0.f, bt1 = 0.f;
    float3 bo = org, bd = dir;
    for (int i = 0; i < (SINGLE ? 1 : nInst); ++i) {
      if (i == last)
        continue;
      const InstanceDev &in = inst[SINGLE ? 0 : i];
      float3 lo = org, ld = dir;
      if (!in.identity) {
        lo = xfmPoint(in.xfm, org);
        ld = xfmVector(in.xfm, dir);
      }
      float t0, t1;
      if (!intersectVolumeBox(in.v.f.boundsLo, in.v.f.boundsHi, lo, ld, rayLower, rayUpper, t0, t1))
        continue;
      if (best < 0 || t0 < bt0) {
        best = i;
        bt0 = t0;
        bt1 = t1;
        bo = lo;
        bd = ld;
      }
    }
    if (best < 0)
      break;
    const InstanceDev &in = inst[SINGLE ? 0 : best];
    bt1 = fminf(tfar, bt1);
    float3 alb = f3(0.f, 0.f, 0.f);
    float ext = 0.f, tr = 0.f;
    float d;
    if (KIND == FIELD_NANOVDB_QUANT || (KIND < 0 && in.v.f.kind == FIELD_NANOVDB_QUANT))
      d = sampleDistanceSegment<FIELD_NANOVDB_QUANT>(in.v, tfOf(SINGLE ? 0 : best), bo, bd, bt0, bt1, rng, alb, ext, tr);
    else if (KIND == FIELD_NANOVDB || (KIND < 0 && in.v.f.kind == FIELD_NANOVDB))
      d = sampleDistanceSegment<FIELD_NANOVDB>(in.v, tfOf(SINGLE ? 0 : best), bo, bd, bt0, bt1, rng, alb, ext, tr);
    else
      d = sampleDistanceSegment<FIELD_STRUCTURED>(in.v, tfOf(SINGLE ? 0 : best), bo, bd, bt0, bt1, rng, alb, ext, tr);
    if (d < depth) {
      depth = d;
      albedo = alb;
      extinction = ext;
      transmittance = tr;
    }
    rayLower = bt1 + 1e-3f;
    last = best;
  }
  return depth;
}

// The raygen loop of DiffusePathTracer_ptx.cu:96-215 without surfaces.  `depth`/`Lw` persist across the
// numIterations loop exactly as PathData does in the reference (declared outside the loop).
struct DptPath
{
  int depth;
  float3 Lw;
};

template <bool SINGLE, int KIND, typename TfSelect>
__device__ __forceinline__ float3 dptTracePath(const InstanceDev *__restrict__ inst, const int nInst, TfSelect tfOf,
    float3 org, float3 dir, const int maxDepth, const float occlusionDistance, const float ambientIntensity,
    const float4 bg, Philox &rng, DptPath &path)
{
  float tmin = 0.f, tmax = FLT_MAX;
  while (true) {
    float3 volumeColor = f3(0.f, 0.f, 0.f);
    float volumeOpacity = 0.f, Tr = 0.f;
    const float volumeDepth = sampleDistanceAllVolumes<SINGLE, KIND>(
        inst, nInst, tfOf, org, dir, tmin, tmax, rng, volumeColor, volumeOpacity, Tr);
    const bool volumeHit = Tr < 1.f;
    if (!volumeHit)
      break;
    if (path.depth++ >= maxDepth) {
      path.Lw = f3(0.f, 0.f, 0.f);
      break;
    }
    const float3 pos = madd3(dir, volumeDepth, org);
    path.Lw = path.Lw * volumeColor;
    const float P = max3(path.Lw); // Russian roulette
    if (P < .2f) {
      if (rng.uniform() > P) {
        path.Lw = f3(0.f, 0.f, 0.f);
        break;
      }
      path.Lw = f3(path.Lw.x / P, path.Lw.y / P, path.Lw.z / P);
    }
    const float3 scatterDir = sampleUnitSphere(rng, f3(-dir.x, -dir.y, -dir.z));
    org = pos;
    dir = scatterDir;
    tmin = 0.f;
    tmax = occlusionDistance;
  }
  return path.depth ? path.Lw * ambientIntensity : f3(bg.x, bg.y, bg.z);
}

} // namespace dvr
