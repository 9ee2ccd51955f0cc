// ref_host.cpp — the reference's OWN host helpers for this path, compiled in place from
// /root/reference (utility/colorMapHelpers.h, gpu/gpu_math.h).  TEST INFRASTRUCTURE (oracle).
// Exposes the transfer-function discretisation so oracle/ and the product's dvr_tf_discretize can
// be checked against the real generateLinearPositions / getInterpolatedValue.  The 20-line loop
// of TransferFunction1D::discritizeTFData (scene/volume/TransferFunction1D.cpp:101-150) is restated
// around them because that class needs helium/ANARI-SDK to compile.
#include "gpu/gpu_math.h"
#include "utility/colorMapHelpers.h"

#include <cstddef>
#include <vector>

using namespace visrtx;

extern "C" int refhost_tf_discretize(const float *color, size_t nColor, int colorChannels, const float *opacity,
    size_t nOpacity, const float uniformColor[4], float uniformOpacity, const float valueRange[2], float *outRgba)
{
  const box1 range(valueRange[0], valueRange[1]);
  const size_t tfDim = 256;
  std::vector<float> cpos, opos;
  Span<float> cPositions, oPositions;
  if (color) {
    cpos = generateLinearPositions(nColor, range);
    cPositions = make_Span(cpos.data(), cpos.size());
  }
  if (opacity) {
    opos = generateLinearPositions(nOpacity, range);
    oPositions = make_Span(opos.data(), opos.size());
  }
  for (size_t i = 0; i < tfDim; i++) {
    const float p = float(i) / (tfDim - 1);
    vec4 c(uniformColor[0], uniformColor[1], uniformColor[2], uniformColor[3]);
    if (color) {
      if (colorChannels == 3)
        c = vec4(getInterpolatedValue((const vec3 *)color, cPositions, range, p), 1.f);
      else
        c = getInterpolatedValue((const vec4 *)color, cPositions, range, p);
    }
    const float o = opacity ? getInterpolatedValue(opacity, oPositions, range, p) : uniformOpacity;
    outRgba[4 * i + 0] = c.x;
    outRgba[4 * i + 1] = c.y;
    outRgba[4 * i + 2] = c.z;
    outRgba[4 * i + 3] = c.w * o;
  }
  return 0;
}

// ray/box helper of the reference (gpu/gpu_math.h intersectBox) for the slab-test cross-check
extern "C" int refhost_position(float v, float lo, float hi, float *out)
{
  *out = position(v, box1(lo, hi));
  return 0;
}

// ---- NanoVDB (vendored with the reference, external/nanovdb 32.7.0) ---------------------------------------
// Synthetic fog-sphere grids for the NanoVDB sampler tests (BASELINE config C5), plus compile-time checks
// of the binary layout the product's own tree reader (visrtx_b200/csrc/dvr_nanovdb.cuh) assumes.
#include <nanovdb/NanoVDB.h>
#include <nanovdb/tools/CreatePrimitives.h>
#include <nanovdb/math/SampleFromVoxels.h>
#include <cstring>

namespace {
using LeafF = nanovdb::NanoLeaf<float>;
using LowerF = nanovdb::NanoLower<float>;
using UpperF = nanovdb::NanoUpper<float>;
using RootF = nanovdb::NanoRoot<float>;
static_assert(sizeof(nanovdb::GridData) == 672, "GridData");
static_assert(sizeof(nanovdb::TreeData) == 64, "TreeData");
static_assert(offsetof(nanovdb::GridData, mMap) == 296, "mMap");
static_assert(offsetof(nanovdb::GridData, mWorldBBox) == 560, "mWorldBBox");
static_assert(offsetof(nanovdb::GridData, mVoxelSize) == 608, "mVoxelSize");
static_assert(offsetof(nanovdb::GridData, mGridType) == 636, "mGridType");
static_assert(offsetof(nanovdb::Map, mInvMatF) == 36 && offsetof(nanovdb::Map, mVecF) == 72, "Map");
static_assert(sizeof(RootF::DataType) == 64, "RootData<float>");
static_assert(sizeof(RootF::DataType::Tile) == 32, "Root tile");
static_assert(sizeof(UpperF::DataType) == 270400, "upper node");
static_assert(sizeof(LowerF::DataType) == 33856, "lower node");
static_assert(sizeof(LeafF::DataType) == 2144, "leaf node");
static_assert(offsetof(UpperF::DataType, mTable) == 8256, "upper table");
static_assert(offsetof(UpperF::DataType, mChildMask) == 32 + 4096, "upper child mask");
static_assert(offsetof(LowerF::DataType, mTable) == 1088, "lower table");
static_assert(offsetof(LowerF::DataType, mChildMask) == 32 + 512, "lower child mask");
static_assert(offsetof(LeafF::DataType, mValues) == 96, "leaf values");
#ifndef NANOVDB_USE_SINGLE_ROOT_KEY
#error "the product's reader assumes 64-bit root keys"
#endif
} // namespace

// returns the byte size; copies min(size, capacity) bytes of the grid buffer into out (may be NULL to query)
extern "C" size_t refhost_nvdb_fog_sphere(double radius, double voxelSize, double halfWidth, const double center[3],
    void *out, size_t capacity)
{
  auto h = nanovdb::tools::createFogVolumeSphere<float>(
      radius, nanovdb::Vec3d(center[0], center[1], center[2]), voxelSize, halfWidth);
  const size_t n = h.size();
  if (out)
    std::memcpy(out, h.data(), n < capacity ? n : capacity);
  return n;
}

// quantised grid types (GridType 13..16): the same sphere through the real CreateNanoGrid quantiser.
// gridType: 1 Float, 13 Fp4, 14 Fp8, 15 Fp16, 16 FpN (tolerance < 0 = NanoVDB's default oracle tolerance)
namespace {
static_assert(sizeof(nanovdb::NanoLeaf<nanovdb::Fp4>::DataType) == 96 + 256, "Fp4 leaf");
static_assert(sizeof(nanovdb::NanoLeaf<nanovdb::Fp8>::DataType) == 96 + 512, "Fp8 leaf");
static_assert(sizeof(nanovdb::NanoLeaf<nanovdb::Fp16>::DataType) == 96 + 1024, "Fp16 leaf");
static_assert(sizeof(nanovdb::NanoLeaf<nanovdb::FpN>::DataType) == 96, "FpN leaf header");
static_assert(offsetof(nanovdb::NanoLeaf<nanovdb::Fp8>::DataType, mMinimum) == 80, "mMinimum");
static_assert(offsetof(nanovdb::NanoLeaf<nanovdb::Fp8>::DataType, mQuantum) == 84, "mQuantum");
static_assert(offsetof(nanovdb::NanoLeaf<nanovdb::Fp8>::DataType, mFlags) == 15, "mFlags");
static_assert(offsetof(nanovdb::NanoLeaf<nanovdb::Fp8>::DataType, mCode) == 96, "mCode");
static_assert(sizeof(nanovdb::NanoUpper<nanovdb::Fp4>::DataType) == 270400, "upper node (Fp)");
static_assert(sizeof(nanovdb::NanoLower<nanovdb::FpN>::DataType) == 33856, "lower node (Fp)");
static_assert(sizeof(nanovdb::NanoRoot<nanovdb::Fp16>::DataType) == 64, "root (Fp)");

template <typename BuildT>
int sampleGrid(const void *gridBlob, const float *xyzWorld, int n, float *out)
{
  const auto *grid = reinterpret_cast<const nanovdb::NanoGrid<BuildT> *>(gridBlob);
  auto acc = grid->getAccessor();
  auto sampler = nanovdb::math::createSampler<1>(acc);
  for (int i = 0; i < n; ++i) {
    const auto loc = nanovdb::Vec3d(xyzWorld[3 * i], xyzWorld[3 * i + 1], xyzWorld[3 * i + 2]);
    out[i] = sampler(nanovdb::math::Vec3d(grid->worldToIndexF(loc)));
  }
  return 0;
}
} // namespace

extern "C" size_t refhost_nvdb_fog_sphere_typed(unsigned gridType, double radius, double voxelSize, double halfWidth,
    const double center[3], float tolerance, void *out, size_t capacity)
{
  const nanovdb::Vec3d c(center[0], center[1], center[2]);
  nanovdb::GridHandle<> h;
  switch (gridType) {
  case 1: h = nanovdb::tools::createFogVolumeSphere<float>(radius, c, voxelSize, halfWidth); break;
  case 13: h = nanovdb::tools::createFogVolumeSphere<nanovdb::Fp4>(radius, c, voxelSize, halfWidth); break;
  case 14: h = nanovdb::tools::createFogVolumeSphere<nanovdb::Fp8>(radius, c, voxelSize, halfWidth); break;
  case 15: h = nanovdb::tools::createFogVolumeSphere<nanovdb::Fp16>(radius, c, voxelSize, halfWidth); break;
  case 16:
    h = nanovdb::tools::createFogVolumeSphere<nanovdb::FpN>(radius, c, voxelSize, halfWidth, nanovdb::Vec3d(0.0),
        "sphere_fog", nanovdb::tools::StatsMode::Default, nanovdb::CheckMode::Default, tolerance, false);
    break;
  default: return 0;
  }
  const size_t n = h.size();
  if (out)
    std::memcpy(out, h.data(), n < capacity ? n : capacity);
  return n;
}

// the reference's own sampler on the host (sampleSpatialField.h:80-109 uses exactly these calls), dispatched on
// the grid type like gpu/volumeIntegration.h:128-159
extern "C" int refhost_nvdb_sample(const void *gridBlob, const float *xyzWorld, int n, float *out)
{
  switch (reinterpret_cast<const nanovdb::GridData *>(gridBlob)->mGridType) {
  case nanovdb::GridType::Float: return sampleGrid<float>(gridBlob, xyzWorld, n, out);
  case nanovdb::GridType::Fp4: return sampleGrid<nanovdb::Fp4>(gridBlob, xyzWorld, n, out);
  case nanovdb::GridType::Fp8: return sampleGrid<nanovdb::Fp8>(gridBlob, xyzWorld, n, out);
  case nanovdb::GridType::Fp16: return sampleGrid<nanovdb::Fp16>(gridBlob, xyzWorld, n, out);
  case nanovdb::GridType::FpN: return sampleGrid<nanovdb::FpN>(gridBlob, xyzWorld, n, out);
  default: return -1;
  }
}

// GridData::isValid + whole-buffer validation the way NvdbRegularField::finalize relies on it
extern "C" int refhost_nvdb_is_valid(const void *gridBlob)
{
  return reinterpret_cast<const nanovdb::GridData *>(gridBlob)->isValid() ? 1 : 0;
}

extern "C" int refhost_nvdb_info(const void *gridBlob, double worldBBox[6], double voxelSize[3], int indexBBox[6],
    unsigned *gridType, unsigned long long *activeVoxels)
{
  const auto *grid = reinterpret_cast<const nanovdb::NanoGrid<float> *>(gridBlob);
  const auto wb = grid->worldBBox();
  const auto ib = grid->indexBBox();
  for (int i = 0; i < 3; ++i) {
    worldBBox[i] = wb.min()[i];
    worldBBox[3 + i] = wb.max()[i];
    voxelSize[i] = grid->voxelSize()[i];
    indexBBox[i] = ib.min()[i];
    indexBBox[3 + i] = ib.max()[i];
  }
  *gridType = (unsigned)grid->gridType();
  *activeVoxels = grid->activeVoxelCount();
  return 0;
}

// ---- .nvdb FILES written by the reference's own NanoVDB I/O (nanovdb::io::writeGrid) -------------------------
// Fixtures for the product's import_NVDB restatement: codec 0 = NONE, 1 = ZIP (zlib).  raw != 0 writes the bare
// grid buffer instead (GridHandle::write), which nanovdb::io::readGrid accepts as well.
#define NANOVDB_USE_ZIP 1
#include <nanovdb/io/IO.h>
#include <nanovdb/tools/GridStats.h>

extern "C" int refhost_nvdb_write_file(const char *path, unsigned gridType, double radius, int codec, int raw)
{
  const nanovdb::Vec3d c(0.0);
  nanovdb::GridHandle<> h;
  switch (gridType) {
  case 1: h = nanovdb::tools::createFogVolumeSphere<float>(radius, c, 1.0, 3.0); break;
  case 13: h = nanovdb::tools::createFogVolumeSphere<nanovdb::Fp4>(radius, c, 1.0, 3.0); break;
  case 14: h = nanovdb::tools::createFogVolumeSphere<nanovdb::Fp8>(radius, c, 1.0, 3.0); break;
  case 15: h = nanovdb::tools::createFogVolumeSphere<nanovdb::Fp16>(radius, c, 1.0, 3.0); break;
  case 16: h = nanovdb::tools::createFogVolumeSphere<nanovdb::FpN>(radius, c, 1.0, 3.0); break;
  default: return -1;
  }
  try {
    if (raw) {
      std::ofstream os(path, std::ios::out | std::ios::binary);
      os.write((const char *)h.data(), h.size());
    } else
      nanovdb::io::writeGrid(path, h, codec == 1 ? nanovdb::io::Codec::ZIP : nanovdb::io::Codec::NONE);
  } catch (const std::exception &) {
    return -2;
  }
  return 0;
}

// what import_NVDB.cpp:25-88 reads back: readGrid + (updateGridStats when the grid has no min/max) + root min/max
extern "C" int refhost_nvdb_read_file(const char *path, void *out, size_t capacity, size_t *size, float minMax[2])
{
  try {
    auto grid = nanovdb::io::readGrid(path);
    auto metadata = grid.gridMetaData();
    minMax[0] = std::numeric_limits<float>::max();
    minMax[1] = std::numeric_limits<float>::lowest();
#define REFHOST_CASE(T)                                                                       \
  {                                                                                           \
    if (!metadata->hasMinMax())                                                               \
      nanovdb::tools::updateGridStats(grid.grid<T>(), nanovdb::tools::StatsMode::MinMax);     \
    minMax[0] = grid.grid<T>()->tree().root().minimum();                                      \
    minMax[1] = grid.grid<T>()->tree().root().maximum();                                      \
    break;                                                                                    \
  }
    switch (metadata->gridType()) {
    case nanovdb::GridType::Fp4: REFHOST_CASE(nanovdb::Fp4)
    case nanovdb::GridType::Fp8: REFHOST_CASE(nanovdb::Fp8)
    case nanovdb::GridType::Fp16: REFHOST_CASE(nanovdb::Fp16)
    case nanovdb::GridType::FpN: REFHOST_CASE(nanovdb::FpN)
    case nanovdb::GridType::Float: REFHOST_CASE(float)
    default: break;
    }
#undef REFHOST_CASE
    *size = grid.size();
    if (out)
      std::memcpy(out, grid.data(), grid.size() < capacity ? grid.size() : capacity);
  } catch (const std::exception &) {
    return -2;
  }
  return 0;
}

// a segment file holding TWO grids (float fog sphere r = radius, then an Fp8 sphere r = radius + 1), written by
// nanovdb::io::writeGrids — import_NVDB reads grid #0 of such files (readGrid(file, n = 0))
extern "C" int refhost_nvdb_write_two_grid_file(const char *path, double radius, int codec)
{
  try {
    std::vector<nanovdb::GridHandle<>> handles;
    handles.push_back(nanovdb::tools::createFogVolumeSphere<float>(radius, nanovdb::Vec3d(0.0), 1.0, 3.0));
    handles.push_back(nanovdb::tools::createFogVolumeSphere<nanovdb::Fp8>(radius + 1.0, nanovdb::Vec3d(0.0), 1.0, 3.0));
    nanovdb::io::writeGrids(path, handles, codec == 1 ? nanovdb::io::Codec::ZIP : nanovdb::io::Codec::NONE);
  } catch (const std::exception &) {
    return -2;
  }
  return 0;
}
