// dvr_march.cuh — the ray-march hot loop (K1) as device functions shared by the frame kernel and
// the sort-last partial kernel.
//
// Reproduces, per pixel-sample, gpu/volumeIntegration.h:64-165,317-350 (fixed-step
// emission-absorption with the reference's double jitter and early termination at 0.99),
// scene/Intersectors_ptx.cu:248-274 (slab test, clamped to the ray interval) and
// gpu/sampleSpatialField.h:54-78 (normalised-coordinate hardware trilinear fetch).
#pragma once

#include "dvr_device.cuh"

namespace dvr {

#ifndef DVR_BATCH
#define DVR_BATCH 4 // field fetches issued back-to-back before compositing (memory-level parallelism)
#endif
#ifndef DVR_BATCH_NVDB
#define DVR_BATCH_NVDB 2 // samples come from apron bricks (8 plain loads): A/B on C5 with 2 CTAs/SM: 1 (3 CTAs) / 2 / 4 = 2450/2909/2958 fps;
                         // the tree-walk fallback alone preferred 1 (810 vs 635 fps at 2)
#endif

#ifndef DVR_FASTPOW
#define DVR_FASTPOW 0
#endif
// pow(1 - alpha, dt/unitDistance), volumeIntegration.h:93
__device__ __forceinline__ float stepPow(float x, float e)
{
#if DVR_FASTPOW == 0
  return powf(x, e);
#elif DVR_FASTPOW == 1
  return exp2f(e * log2f(x));
#else
  return __powf(x, e);
#endif
}

struct MarchStats
{
  unsigned long long taken;
  unsigned long long skipped;
};

// object-space ray/box slab test; returns false when the box is missed or outside [tmin,tmax]
__device__ __forceinline__ bool intersectVolumeBox(
    const float3 lo, const float3 hi, const float3 org, const float3 dir, float tmin, float tmax,
    float &t0, float &t1)
{
  const float3 inv = f3(__fdiv_rn(1.f, dir.x), __fdiv_rn(1.f, dir.y), __fdiv_rn(1.f, dir.z));
  const float3 mins = (lo - org) * inv;
  const float3 maxs = (hi - org) * inv;
  const float3 nears = f3(fminf(mins.x, maxs.x), fminf(mins.y, maxs.y), fminf(mins.z, maxs.z));
  const float3 fars = f3(fmaxf(mins.x, maxs.x), fmaxf(mins.y, maxs.y), fmaxf(mins.z, maxs.z));
  const float tn = max3(nears);
  const float tf = min3(fars);
  if (!(tn < tf))
    return false;
  // OptiX only runs the intersection program for boxes that overlap the ray interval
  if (tf < tmin || tn > tmax)
    return false;
  t0 = fmaxf(tmin, fminf(tn, tmax));
  t1 = fmaxf(tmin, fminf(tf, tmax));
  return true;
}

__device__ __forceinline__ float3 xfmPoint(const float *m, float3 p)
{
  return f3(__fmaf_rn(m[2], p.z, __fmaf_rn(m[1], p.y, __fmaf_rn(m[0], p.x, m[3]))),
      __fmaf_rn(m[6], p.z, __fmaf_rn(m[5], p.y, __fmaf_rn(m[4], p.x, m[7]))),
      __fmaf_rn(m[10], p.z, __fmaf_rn(m[9], p.y, __fmaf_rn(m[8], p.x, m[11]))));
}
__device__ __forceinline__ float3 xfmVector(const float *m, float3 v)
{
  return f3(__fmaf_rn(m[2], v.z, __fmaf_rn(m[1], v.y, __fmul_rn(m[0], v.x))),
      __fmaf_rn(m[6], v.z, __fmaf_rn(m[5], v.y, __fmul_rn(m[4], v.x))),
      __fmaf_rn(m[10], v.z, __fmaf_rn(m[9], v.y, __fmul_rn(m[8], v.x))));
}

// texture coordinates of an object-space position: sampleSpatialField.h:66-70
__device__ __forceinline__ float3 fieldTexCoord(const FieldDev &f, const float3 halfSpacing, float3 p)
{
  return ((p - f.origin) + halfSpacing) * f.invSpacing;
}

template <bool SLAB>
__device__ __forceinline__ float fieldFetch(const FieldDev &f, float3 tc)
{
  if (SLAB) {
    // remap the global normalised z onto the resident slices (see DESIGN.md "sort-last slabs")
    const float zb = tc.z * (float)f.dims.z - (float)f.zTexBegin;
    tc.z = zb / (float)f.texDepth;
  }
  return tex3D<float>(f.tex, tc.x, tc.y, tc.z);
}

// Sampling coordinate of an object-space position: normalised texture coordinate (structuredRegular) or
// index-space coordinate (NanoVDB); sampleSpatialField.h:66-70 / :91-95.
template <int KIND>
__device__ __forceinline__ float3 fieldCoord(const FieldDev &f, const float3 halfSpacing, float3 p)
{
  if (KIND >= FIELD_NANOVDB)
    return nvdbWorldToIndex(f.nv, p);
  return fieldTexCoord(f, halfSpacing, p);
}

// continuous voxel coordinate whose floor is the lower tap index (macrocell / slab bookkeeping only)
template <int KIND>
__device__ __forceinline__ float3 coordToVoxel(const FieldDev &f, float3 c)
{
  if (KIND >= FIELD_NANOVDB)
    return f3(c.x - (float)f.nv.bboxMin.x, c.y - (float)f.nv.bboxMin.y, c.z - (float)f.nv.bboxMin.z);
  return f3(c.x * (float)f.dims.x - 0.5f, c.y * (float)f.dims.y - 0.5f, c.z * (float)f.dims.z - 0.5f);
}

template <int KIND, bool SLAB>
__device__ __forceinline__ float fieldSample(const FieldDev &f, NvdbCache &cache, float3 c)
{
  if (KIND == FIELD_NANOVDB_QUANT)
    return nvdbSampleTrilinear<true>(f.nv, cache, c);
  if (KIND == FIELD_NANOVDB)
    return nvdbSampleTrilinear<false>(f.nv, cache, c);
  return fieldFetch<SLAB>(f, c);
}

// `while (n > 0 && t <= tUpper) { t += step; --n; }` — the reference's lattice is DEFINED by repeated float
// addition, so skipped lattice points must land on exactly the values that loop produces.  Inside one binade
// [2^e, 2^(e+1)) every float is a multiple of the same ulp u, so each round-to-nearest add moves t by the same
// amount inc = fl(t + step) - t (a multiple of u) unless t + step falls exactly half-way between two floats
// (ties-to-even alternates).  Hence k further adds give t + k*inc exactly as long as the results stay in the
// binade: one real add measures inc, the rest of the binade is jumped in closed form (exact in double), and the
// add that crosses into the next binade is again a real one.  O(binades crossed) instead of O(n); bit-identical
// to the loop (tests: skipping on/off, sort-last).
__device__ __forceinline__ float latticeAdvance(float t, const float step, int n, const float tUpper, int &taken)
{
  taken = 0;
  while (n > 0 && t <= tUpper) {
    const float t0 = t;
    t = __fadd_rn(t0, step); // a real step
    ++taken;
    --n;
    if (n == 0 || !(t <= tUpper))
      break;
    const float inc = __fsub_rn(t, t0); // exact: both are multiples of the binade's ulp
    const int e0 = (__float_as_int(t0) >> 23) & 0xff, e1 = (__float_as_int(t) >> 23) & 0xff;
    const int es = (__float_as_int(step) >> 23) & 0xff;
    if (!(t0 > 0.f) || e0 != e1 || e1 == 0 || e1 == 0xff || e1 - es > 28 || es - e1 > 1)
      continue; // crossing a binade (or degenerate operands): keep taking real steps
    if (!(inc > 0.f)) { // step is below half an ulp of t: the loop would spin in place for all remaining adds
      taken += n;
      n = 0;
      break;
    }
    const double u = __longlong_as_double((long long)(e1 - 127 - 23 + 1023) << 52); // ulp of the binade
    const double err = ((double)t0 + (double)step) - (double)t;                       // exact
    if (fabs(err) * 2.0 == u)
      continue; // a tie: round-to-even makes the increment alternate
    // Step counts are estimated in fp32 (one reciprocal) and then made exact with double compares; an estimate
    // beyond the n adds still wanted needs no correction.
    const double td = (double)t, incd = (double)inc;
    const double top = __longlong_as_double((long long)(e1 - 127 + 1 + 1023) << 52); // 2^(e+1)
    const float rinc = __frcp_rn(inc);
    const float lim = (float)n + 2.f;
    // kb: adds whose RESULT stays below 2^(e+1)
    long long kb = (long long)fminf(floorf(__fmul_rn((float)(top - td), rinc)), lim);
    if (kb <= (long long)n + 1) {
      while (kb > 0 && td + (double)kb * incd >= top)
        --kb;
      while (td + (double)(kb + 1) * incd < top)
        ++kb;
    }
    // ku: adds the loop executes before t > tUpper
    long long ku = (long long)fminf(floorf(__fmul_rn((float)((double)tUpper - td), rinc)) + 1.f, lim);
    if (ku <= (long long)n + 1) {
      while (ku > 1 && td + (double)(ku - 1) * incd > (double)tUpper)
        --ku;
      while (td + (double)ku * incd <= (double)tUpper)
        ++ku;
    }
    long long k = kb < ku ? kb : ku;
    if (k > (long long)n)
      k = n;
    if (k > 0) {
      t = (float)(td + (double)k * incd); // exact: a multiple of u below 2^(e+1)
      taken += (int)k;
      n -= (int)k;
    }
  }
  return t;
}

// One volume segment [tLower(after jitter #1), tUpper] of one ray.
//   SKIP : consult the per-macrocell majorants and step over fully transparent cells on the
//          SAME sample lattice (t advances by repeated `t += step`, so the taken samples are
//          bit-identical to the unskipped march)
//   SLAB : only samples whose cell slice lies in [zOwnBegin,zOwnEnd) are taken (sort-last)
//   STATS: count samples
//   G    : depth lanes.  G > 1: G adjacent lanes of the warp (lane % G = depth slot) march the SAME ray; every
//          lane carries the ray's full state redundantly (bit-identical arithmetic), lane g fetches and classifies
//          the lattice points j with j % G == g of each iteration, and all G lanes then run the reference's
//          sequential front-to-back composite over the iteration's BATCH*G samples in lattice order, reading each
//          classified sample from its owner with a sub-warp shuffle.  The image is bit-identical to G == 1; the
//          dependent-latency chain of a ray is G times shorter and adjacent lanes fetch adjacent voxels.
template <bool SKIP, bool SLAB, bool STATS, int KIND, int G = 1>
__device__ __forceinline__ void marchSegment(const VolumeDev &v, const float4 *__restrict__ tf,
    const float3 org, const float3 dir, float t, const float tUpper, const float invSamplingRate,
    Philox &rng, float3 &color, float &opacity, MarchStats &stats, unsigned int *cellBitmap)
{
  constexpr int BATCH = KIND >= FIELD_NANOVDB ? DVR_BATCH_NVDB : DVR_BATCH;
  constexpr int NS = BATCH * G; // lattice points per iteration of this ray
  static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "depth lanes must divide the warp");
  const int g = G > 1 ? (int)(threadIdx.x & (G - 1)) : 0;
  const unsigned gmask = G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31u) - (unsigned)g));
  const FieldDev &f = v.f;
  const float stepSize = __fmul_rn(f.stepSize, invSamplingRate);
  const float exponent = __fmul_rn(stepSize, v.oneOverUnitDistance);
  t = __fmaf_rn(stepSize, rng.uniform(), t); // jitter #2, volumeIntegration.h:83

  const float3 halfSpacing = 0.5f * f.spacing;
  const float vrLo = v.vrLower, vrHi = v.vrUpper;
  const float invRange = __fdiv_rn(1.0f, __fsub_rn(vrHi, vrLo)); // gpu_math.h:176-180, hoisted out of the loop
  float transmittance = 1.f;

  // d(voxel coordinate)/dt, used by SKIP/SLAB bookkeeping only (never for the sample position)
  float3 dvox;
  if (KIND >= FIELD_NANOVDB)
    dvox = f3(dir.x * f.nv.invMat[0] + dir.y * f.nv.invMat[1] + dir.z * f.nv.invMat[2],
        dir.x * f.nv.invMat[3] + dir.y * f.nv.invMat[4] + dir.z * f.nv.invMat[5],
        dir.x * f.nv.invMat[6] + dir.y * f.nv.invMat[7] + dir.z * f.nv.invMat[8]);
  else
    dvox = f3(dir.x * f.invSpacing.x * (float)f.dims.x, dir.y * f.invSpacing.y * (float)f.dims.y,
        dir.z * f.invSpacing.z * (float)f.dims.z);
  const float3 invDvox = f3(dvox.x != 0.f ? 1.f / dvox.x : 0.f, dvox.y != 0.f ? 1.f / dvox.y : 0.f,
      dvox.z != 0.f ? 1.f / dvox.z : 0.f);
  const float invStep = 1.f / stepSize;
  NvdbCache nvCache;
  if (KIND >= FIELD_NANOVDB)
    nvCache.reset();

  if (SLAB) {
    // fast-forward to just before the ray enters the owned z range (sequential adds keep the
    // lattice identical to the single-GPU march)
    const float3 tc0 = fieldTexCoord(f, halfSpacing, madd3(dir, t, org));
    const float z0 = tc0.z * (float)f.dims.z - 0.5f;
    float tEnter = t;
    if (dvox.z > 0.f)
      tEnter = t + ((float)f.zOwnBegin - z0) / dvox.z;
    else if (dvox.z < 0.f)
      tEnter = t + ((float)f.zOwnEnd - z0) / dvox.z;
    tEnter -= 2.f * stepSize;
    // `while (t < tEnter && t <= tUpper) t += step`, in closed form (t < x  <=>  t <= the float just below x)
    if (t < tEnter && t <= tUpper && tEnter == tEnter) {
      const float below = __int_as_float(__float_as_int(tEnter) + (tEnter > 0.f ? -1 : (tEnter < 0.f ? 1 : 0)));
      const float bound = fminf(tUpper, tEnter == 0.f ? -FLT_MIN : below);
      int taken;
      t = latticeAdvance(t, stepSize, 0x7fffffff, bound, taken);
      if (STATS && g == 0)
        stats.skipped += (unsigned long long)taken;
    }
  }

  while (opacity < 0.99f && t <= tUpper) {
    if (SKIP) {
      const float3 xb = coordToVoxel<KIND>(f, fieldCoord<KIND>(f, halfSpacing, madd3(dir, t, org)));
      const int cx = min(max((int)floorf(xb.x), 0), f.dims.x - 1) >> 4;
      const int cy = min(max((int)floorf(xb.y), 0), f.dims.y - 1) >> 4;
      const int cz = min(max((int)floorf(xb.z), 0), f.dims.z - 1) >> 4;
      // three-level test: the 256^3- and 64^3-voxel blocks first (long empty runs in one hop), then the 16^3
      // macrocell; the loads are issued together
      const float majorant = __ldg(&v.maxOpacities[(size_t)cz * f.gridDims.x * f.gridDims.y
          + (size_t)cy * f.gridDims.x + cx]);
      const float coarse = __ldg(&v.maxOpacitiesCoarse[(size_t)(cz >> 2) * v.coarseDims.x * v.coarseDims.y
          + (size_t)(cy >> 2) * v.coarseDims.x + (cx >> 2)]);
      const float coarse2 = __ldg(&v.maxOpacitiesCoarse2[(size_t)(cz >> 4) * v.coarse2Dims.x * v.coarse2Dims.y
          + (size_t)(cy >> 4) * v.coarse2Dims.x + (cx >> 4)]);
      if (majorant <= 0.f) {
        // distance (in t) to the nearest face of the empty region along the ray
        const int lvl = coarse2 <= 0.f ? 4 : (coarse <= 0.f ? 2 : 0); // cell-index shift of the empty region
        const int sh = 4 + lvl;
        const int rx = cx >> lvl, ry = cy >> lvl, rz = cz >> lvl;
        const float bx = dvox.x > 0.f ? (float)((rx + 1) << sh) : (float)(rx << sh);
        const float by = dvox.y > 0.f ? (float)((ry + 1) << sh) : (float)(ry << sh);
        const float bz = dvox.z > 0.f ? (float)((rz + 1) << sh) : (float)(rz << sh);
        // (reciprocals hoisted out of the loop: dt only sizes the hop, one whole step of margin absorbs its rounding)
        const float ex = dvox.x != 0.f ? (bx - xb.x) * invDvox.x : FLT_MAX;
        const float ey = dvox.y != 0.f ? (by - xb.y) * invDvox.y : FLT_MAX;
        const float ez = dvox.z != 0.f ? (bz - xb.z) * invDvox.z : FLT_MAX;
        const float dt = fminf(fminf(ex, ey), ez);
        // Whole steps that stay strictly inside the cell (one step of safety margin); at least THIS lattice
        // point is skippable on its own: the cell containing its lower tap has majorant 0, so its fetch
        // would classify to alpha == 0 exactly and contribute nothing.
        const int n = max((int)floorf(fminf(dt * invStep, 1.0e6f)) - 1, 1);
        int taken;
        t = latticeAdvance(t, stepSize, n, tUpper, taken);
        if (STATS && g == 0)
          stats.skipped += (unsigned long long)taken;
        continue;
      }
    }

    float ts[BATCH];
    float s[BATCH];
    float tt = t;
#pragma unroll
    for (int j = 0; j < NS; ++j) { // the lattice by repeated addition, exactly the reference's `t += step`
      if (G == 1 || (j % G) == g)
        ts[j / G] = tt;
      tt = __fadd_rn(tt, stepSize);
    }
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      s[k] = __int_as_float(0x7fc00000);
      if (ts[k] <= tUpper) {
        const float3 p = madd3(dir, ts[k], org);
        const float3 tc = fieldCoord<KIND>(f, halfSpacing, p);
        bool own = true;
        if (SLAB) {
          const int zc = min(max((int)floorf(tc.z * (float)f.dims.z - 0.5f), 0), f.dims.z - 1);
          own = zc >= f.zOwnBegin && zc < f.zOwnEnd;
        }
        if (own) {
          s[k] = fieldSample<KIND, SLAB>(f, nvCache, tc);
          if (STATS) {
            stats.taken++;
            if (cellBitmap) {
              const float3 xv = coordToVoxel<KIND>(f, tc);
              const int cx = min(max((int)floorf(xv.x), 0), f.dims.x - 1) >> 4;
              const int cy = min(max((int)floorf(xv.y), 0), f.dims.y - 1) >> 4;
              const int cz = min(max((int)floorf(xv.z), 0), f.dims.z - 1) >> 4;
              const size_t c = (size_t)cz * f.gridDims.x * f.gridDims.y + (size_t)cy * f.gridDims.x + cx;
              atomicOr(&cellBitmap[c >> 5], 1u << (c & 31));
            }
          }
        }
      }
    }
    // classify + opacity correction for the whole batch first (independent chains => ILP), then the
    // short sequential front-to-back composite with the reference's per-sample termination test
    float4 co[BATCH];
    float st[BATCH];
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const float c = __fmul_rn(__fsub_rn(fmaxf(vrLo, fminf(s[k], vrHi)), vrLo), invRange); // position(s, range)
      co[k] = tfLookup(tf, c);
      st[k] = stepPow(__fsub_rn(1.f, co[k].w), exponent);
      // s[k] is NaN for lattice points past the segment / not owned / NaN voxels: skipped like the reference
      if (isnan(s[k]))
        st[k] = s[k];
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int k = j / G;
      float cr = co[k].x, cg = co[k].y, cb = co[k].z, stj = st[k];
      if (G > 1) { // sample j lives in depth lane j % G
        cr = __shfl_sync(gmask, cr, j % G, G);
        cg = __shfl_sync(gmask, cg, j % G, G);
        cb = __shfl_sync(gmask, cb, j % G, G);
        stj = __shfl_sync(gmask, stj, j % G, G);
      }
      if (opacity < 0.99f && !isnan(stj)) {
        const float w = __fmul_rn(transmittance, __fsub_rn(1.f, stj));
        color.x = __fmaf_rn(w, cr, color.x);
        color.y = __fmaf_rn(w, cg, color.y);
        color.z = __fmaf_rn(w, cb, color.z);
        opacity = __fadd_rn(opacity, w);
        transmittance = __fmul_rn(transmittance, stj);
      }
    }
    t = tt;

    if (SLAB) {
      // past the owned range for good?
      const float3 tc = fieldTexCoord(f, halfSpacing, madd3(dir, t, org));
      const float z = tc.z * (float)f.dims.z - 0.5f;
      if ((dvox.z > 0.f && z > (float)f.zOwnEnd + 2.f) || (dvox.z < 0.f && z < (float)f.zOwnBegin - 2.f))
        break;
    }
  }
}

// rayMarchAllVolumes, volumeIntegration.h:317-350, with the OptiX volume-BVH trace replaced by a
// loop over the flattened instance list (closest clamped entry first, the volume marched last
// is excluded from the next search exactly like the lastVolID/lastInstID test of
// Intersectors_ptx.cu:250-252).
// SINGLE: exactly one instance => every access uses the constant index 0, which keeps the texture
// handle and field constants warp-uniform (no divergent-handle loop around the TEX instruction).
// KIND: field kind known at compile time (single-volume kernels), or -1 = decide per instance.
template <bool SKIP, bool SLAB, bool STATS, bool SINGLE, int KIND, int G = 1, typename TfSelect>
__device__ __forceinline__ float rayMarchAllVolumes(const InstanceDev *__restrict__ inst, const int nInst,
    TfSelect tfOf, const float3 org, const float3 dir, const float tfar, const float invSamplingRate,
    Philox &rng, float3 &color, float &opacity, uint32_t &objID, uint32_t &instID, MarchStats &stats,
    unsigned int *cellBitmap, bool &anyHit, const float tnear = 0.f)
{
  float rayLower = tnear; // ray.t.lower of the caller: 0 for a fresh primary ray, past the last surface otherwise
  const float rayUpper = tfar;
  float depth = tfar;
  bool firstHit = true;
  int last = -1;

  do {
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
    if (firstHit) {
      objID = in.v.id;
      instID = in.instId;
      firstHit = false;
      anyHit = true;
    }
    depth = fminf(depth, bt0);
    bt1 = fminf(tfar, bt1);
    // detail::rayMarchVolume: jitter #1 uses the UNSCALED step (volumeIntegration.h:117-120)
    const float tStart = __fmaf_rn(in.v.f.stepSize, rng.uniform(), bt0);
    if (KIND == FIELD_NANOVDB_QUANT || (KIND < 0 && in.v.f.kind == FIELD_NANOVDB_QUANT))
      marchSegment<SKIP, false, STATS, FIELD_NANOVDB_QUANT, G>(
          in.v, tfOf(SINGLE ? 0 : best), bo, bd, tStart, bt1, invSamplingRate, rng, color, opacity, stats, cellBitmap);
    else if (KIND == FIELD_NANOVDB || (KIND < 0 && in.v.f.kind == FIELD_NANOVDB))
      marchSegment<SKIP, false, STATS, FIELD_NANOVDB, G>(
          in.v, tfOf(SINGLE ? 0 : best), bo, bd, tStart, bt1, invSamplingRate, rng, color, opacity, stats, cellBitmap);
    else
      marchSegment<SKIP, SLAB, STATS, FIELD_STRUCTURED, G>(
          in.v, tfOf(SINGLE ? 0 : best), bo, bd, tStart, bt1, invSamplingRate, rng, color, opacity, stats, cellBitmap);
    rayLower = __fadd_rn(bt1, 1e-3f);
    last = best;
  } while (opacity < 0.99f);

  return depth;
}

} // namespace dvr
