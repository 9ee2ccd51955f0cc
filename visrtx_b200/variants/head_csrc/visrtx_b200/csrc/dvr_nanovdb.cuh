// dvr_nanovdb.cuh — device-side reader + trilinear sampler for NanoVDB float grids (BASELINE config C5) and
// the quantised grid types Fp4 / Fp8 / Fp16 / FpN (gpu/volumeIntegration.h:128-159 dispatches the same five).
//
// Replaces SpatialFieldSampler<nanovdb::Grid<NanoTree<float>>> (gpu/sampleSpatialField.h:80-109), which
// calls grid->worldToIndexF() and nanovdb::math::SampleFromVoxels<Accessor,1> of the NanoVDB 32.7.0
// headers vendored with the reference.  The product does not include those headers: the tree walk below
// is written against the published binary layout (GridData 672 B, TreeData 64 B, RootData<float> 64 B +
// 32-byte tiles with 64-bit keys, upper 32^3 / lower 16^3 internal nodes, 8^3 float leaves); the layout
// constants are verified against the real headers by static_asserts in the test infrastructure's
// reference-host helper and the sampler against the reference's own host sampler (tests/test_gpu_nanovdb.py).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dvr {

// byte offsets of the layout (see header comment)
enum : uint32_t
{
  kNvdbRootTiles = 64,
  kNvdbTileSize = 32,
  kNvdbUpperChildMask = 32 + 4096,
  kNvdbUpperTable = 8256,
  kNvdbLowerChildMask = 32 + 512,
  kNvdbLowerTable = 1088,
  kNvdbLeafValues = 96
};

struct NvdbDev
{
  const uint8_t *root; // RootData<float>
  uint32_t tileCount;
  float background;
  float invMat[9]; // Map::mInvMatF
  float vec[3];    // Map::mVecF
  int3 bboxMin;    // index-space bounding box of the active values (RootData::mBBox)
  int codecLog2Bits; // quantised grids: log2(bits per code) — Fp4 2, Fp8 3, Fp16 4, FpN -1 (per leaf: mFlags >> 5)
};

// per-thread cache of the last visited lower node and leaf (the 8 taps of a trilinear stencil almost
// always share them) — the role nanovdb::ReadAccessor plays in the reference
struct NvdbCache
{
  int lx, ly, lz;          // leaf key   (coords >> 3)
  const uint8_t *leaf;     // nullptr: (lx,ly,lz) is a constant region of value leafTile
  float leafTile;
  int nx, ny, nz;          // lower-node key (coords >> 7)
  const uint8_t *lower;
  float qMin, qQuantum;    // quantised grids: LeafFnBase::mMinimum / mQuantum of the cached leaf
  int qLog2Bits;           //                  log2(bits per code) of the cached leaf
  __device__ __forceinline__ void reset()
  {
    lx = ly = lz = nx = ny = nz = 0x7fffffff;
    leaf = lower = nullptr;
    leafTile = 0.f;
    qMin = qQuantum = 0.f;
    qLog2Bits = 0;
  }
};

// Quantised leaves: LeafFnBase (96 B: bbox 16 B with mFlags at byte 15, value mask 64 B, float mMinimum @80,
// float mQuantum @84, 4 x u16 stats) followed by the packed codes.  The header of the cached leaf is read once.
__device__ __forceinline__ void nvdbCacheLeafHeader(const NvdbDev &g, NvdbCache &c)
{
  const float2 mq = *reinterpret_cast<const float2 *>(c.leaf + 80);
  c.qMin = mq.x;
  c.qQuantum = mq.y;
  c.qLog2Bits = g.codecLog2Bits >= 0 ? g.codecLog2Bits : (int)(c.leaf[15] >> 5);
}

// Leaf value n (0..511) of the cached leaf.  LeafData<Fp4|Fp8|Fp16|FpN>::getValue: code * mQuantum + mMinimum
// (one fused multiply-add on the GPU); float leaves: mValues[n].
template <bool QUANT>
__device__ __forceinline__ float nvdbLeafValue(const NvdbCache &c, uint32_t n)
{
  if (QUANT) {
    const int b = c.qLog2Bits;
    uint32_t code = reinterpret_cast<const uint32_t *>(c.leaf + kNvdbLeafValues)[n >> (5 - b)];
    code >>= (n & ((32u >> b) - 1u)) << b;
    code &= (1u << (1u << b)) - 1u;
    return __fmaf_rn((float)code, c.qQuantum, c.qMin);
  }
  return *reinterpret_cast<const float *>(c.leaf + kNvdbLeafValues + 4u * n);
}

__device__ __forceinline__ bool nvdbMaskOn(const uint8_t *mask, uint32_t n)
{
  const uint64_t w = *reinterpret_cast<const uint64_t *>(mask + ((n >> 6) << 3));
  return (w >> (n & 63u)) & 1ull;
}

// Tree::getValue(ijk): value of the voxel regardless of its active state (inactive => tile / background value)
template <bool QUANT = false>
__device__ __forceinline__ float nvdbGetValue(const NvdbDev &g, NvdbCache &c, int x, int y, int z)
{
  const int kx = x >> 3, ky = y >> 3, kz = z >> 3;
  if (kx == c.lx && ky == c.ly && kz == c.lz) {
    if (c.leaf)
      return nvdbLeafValue<QUANT>(c, (uint32_t)(((x & 7) << 6) | ((y & 7) << 3) | (z & 7)));
    return c.leafTile;
  }
  const uint8_t *lower = nullptr;
  if ((x >> 7) == c.nx && (y >> 7) == c.ny && (z >> 7) == c.nz && c.lower)
    lower = c.lower;
  else {
    // root: linear probe of the tile table by key (RootData::probeTile)
    const uint64_t key = (uint64_t)((uint32_t)z >> 12) | ((uint64_t)((uint32_t)y >> 12) << 21)
        | ((uint64_t)((uint32_t)x >> 12) << 42);
    const uint8_t *tile = nullptr;
    for (uint32_t i = 0; i < g.tileCount; ++i) {
      const uint8_t *t = g.root + kNvdbRootTiles + (size_t)i * kNvdbTileSize;
      if (*reinterpret_cast<const uint64_t *>(t) == key) {
        tile = t;
        break;
      }
    }
    if (!tile)
      return g.background;
    const int64_t child = *reinterpret_cast<const int64_t *>(tile + 8);
    if (child == 0)
      return *reinterpret_cast<const float *>(tile + 20);
    const uint8_t *upper = g.root + child;
    const uint32_t n = (uint32_t)((((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7));
    const uint8_t *entry = upper + kNvdbUpperTable + 8u * n;
    if (!nvdbMaskOn(upper + kNvdbUpperChildMask, n))
      return *reinterpret_cast<const float *>(entry);
    lower = upper + *reinterpret_cast<const int64_t *>(entry);
    c.nx = x >> 7;
    c.ny = y >> 7;
    c.nz = z >> 7;
    c.lower = lower;
  }
  const uint32_t n = (uint32_t)((((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3));
  const uint8_t *entry = lower + kNvdbLowerTable + 8u * n;
  c.lx = kx;
  c.ly = ky;
  c.lz = kz;
  if (!nvdbMaskOn(lower + kNvdbLowerChildMask, n)) {
    c.leaf = nullptr;
    c.leafTile = *reinterpret_cast<const float *>(entry);
    return c.leafTile;
  }
  c.leaf = lower + *reinterpret_cast<const int64_t *>(entry);
  if (QUANT)
    nvdbCacheLeafHeader(g, c);
  return nvdbLeafValue<QUANT>(c, (uint32_t)(((x & 7) << 6) | ((y & 7) << 3) | (z & 7)));
}

// grid->worldToIndexF(Vec3d(location)) : Map::applyInverseMapF -> math::matMult(const float*, Vec3d)
// (subtraction in double, operands cast to float, two fmaf + one multiply per row)
__device__ __forceinline__ float3 nvdbWorldToIndex(const NvdbDev &g, float3 p)
{
  const float fx = (float)((double)p.x - (double)g.vec[0]);
  const float fy = (float)((double)p.y - (double)g.vec[1]);
  const float fz = (float)((double)p.z - (double)g.vec[2]);
  return make_float3(__fmaf_rn(fx, g.invMat[0], __fmaf_rn(fy, g.invMat[1], __fmul_rn(fz, g.invMat[2]))),
      __fmaf_rn(fx, g.invMat[3], __fmaf_rn(fy, g.invMat[4], __fmul_rn(fz, g.invMat[5]))),
      __fmaf_rn(fx, g.invMat[6], __fmaf_rn(fy, g.invMat[7], __fmul_rn(fz, g.invMat[8]))));
}

// SampleFromVoxels<Acc,1>::operator()(Vec3d): ijk = floor(xyz), uvw = xyz - ijk, stencil of 8 getValue
// calls, nested lerps a + w*(b - a) with z innermost (math/SampleFromVoxels.h:201-242)
template <bool QUANT = false>
__device__ __forceinline__ float nvdbSampleTrilinear(const NvdbDev &g, NvdbCache &c, float3 idx)
{
  const float fi = floorf(idx.x), fj = floorf(idx.y), fk = floorf(idx.z);
  const int i = (int)fi, j = (int)fj, k = (int)fk;
  const float u = __fsub_rn(idx.x, fi), v = __fsub_rn(idx.y, fj), w = __fsub_rn(idx.z, fk);
  float v000, v001, v011, v010, v100, v101, v111, v110;
  v000 = nvdbGetValue<QUANT>(g, c, i, j, k); // positions the leaf cache on the stencil's base voxel
  if (((i & 7) < 7) & ((j & 7) < 7) & ((k & 7) < 7)) {
    // whole stencil inside the cached leaf (or constant region): seven independent loads, no tree walk
    if (c.leaf) {
      const uint32_t n = (uint32_t)(((i & 7) << 6) | ((j & 7) << 3) | (k & 7));
      v001 = nvdbLeafValue<QUANT>(c, n + 1);
      v010 = nvdbLeafValue<QUANT>(c, n + 8);
      v011 = nvdbLeafValue<QUANT>(c, n + 9);
      v100 = nvdbLeafValue<QUANT>(c, n + 64);
      v101 = nvdbLeafValue<QUANT>(c, n + 65);
      v110 = nvdbLeafValue<QUANT>(c, n + 72);
      v111 = nvdbLeafValue<QUANT>(c, n + 73);
    } else
      v001 = v010 = v011 = v100 = v101 = v110 = v111 = c.leafTile;
  } else {
    v001 = nvdbGetValue<QUANT>(g, c, i, j, k + 1);
    v011 = nvdbGetValue<QUANT>(g, c, i, j + 1, k + 1);
    v010 = nvdbGetValue<QUANT>(g, c, i, j + 1, k);
    v100 = nvdbGetValue<QUANT>(g, c, i + 1, j, k);
    v101 = nvdbGetValue<QUANT>(g, c, i + 1, j, k + 1);
    v111 = nvdbGetValue<QUANT>(g, c, i + 1, j + 1, k + 1);
    v110 = nvdbGetValue<QUANT>(g, c, i + 1, j + 1, k);
  }
#define DVR_LERP(a, b, t) __fmaf_rn((t), __fsub_rn((b), (a)), (a))
  const float r = DVR_LERP(DVR_LERP(DVR_LERP(v000, v001, w), DVR_LERP(v010, v011, w), v),
      DVR_LERP(DVR_LERP(v100, v101, w), DVR_LERP(v110, v111, w), v), u);
#undef DVR_LERP
  return r;
}

} // namespace dvr
