// dvr_internal.h — host-side declarations shared by the translation units of libdvr_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/dvr_b200.h"
#include "dvr_device.cuh"

namespace dvr {

constexpr int kMaxInlineInstances = 4;
#ifndef DVR_TILE_W
#define DVR_TILE_W 8 // one warp renders an 8x4 pixel tile
#endif
constexpr int kTileW = DVR_TILE_W;
constexpr int kTileH = 32 / DVR_TILE_W;
constexpr int kBlockThreads = 256;
constexpr int kMaxSlabs = 16;

// cross-GPU flags of one launch (DvrPeerSync of the C-ABI)
struct SyncDev
{
  uint32_t nSignal, signalValue;
  unsigned int *signal[kMaxSlabs];
  uint32_t nWait, waitValue;
  const unsigned int *wait;
  unsigned int *errorFlag;
};


// kernel parameter block of the frame kernel (passed by value as __grid_constant__)
struct FrameLaunch
{
  uint32_t width, height;
  float invW, invH;
  int format, integrator, frameID, checkerboardID, numIterations;
  float invSamplingRate;
  float4 background;
  cudaTextureObject_t bgTex; // background image (0 = the constant colour)
  int maxDepth;
  float ambientIntensity, occlusionDistance;
  uint32_t tileRank, tileRanks, tileBand;
  uint32_t launchW, launchH; // pixel-sample grid actually launched (half size when checkerboarding)
  uint32_t tilesX, tilesY;
  // Pixels outside [missX0,missX1) x [missY0,missY1) cannot hit any volume (conservative screen rectangle of all
  // instance bounds): they take the background without generating a ray.  missValid == 0: no such knowledge.
  int missValid, missX0, missY0, missX1, missY1;
  // tiles actually scheduled: the tile-aligned rectangle above when sweepOutside is set (the pixels outside it are
  // then handled by dvrBackgroundSweepKernel), else the whole launch grid
  uint32_t tileX0, tileY0, tilesW, tilesH;
  int sweepOutside;
  CameraDev cam;
  BuffersDev fb;
  int nInst;
  InstanceDev inl[kMaxInlineInstances];
  const InstanceDev *ext; // used when nInst > kMaxInlineInstances
  unsigned int *sched;    // [0] next tile, [1] warps finished
  DvrRenderStats *stats;
  unsigned int *cellBitmap;
};

struct PartialLaunch
{
  uint32_t width, height;
  float invW, invH;
  int integrator, frameID;
  float invSamplingRate;
  uint32_t tilesX, tilesY;
  uint32_t tileX0, tileY0, tilesW, tilesH; // tile window actually rendered (partialCullToBounds), else the whole frame
  CameraDev cam;
  InstanceDev inst;
  float4 *partialRgba;
  float *partialDepth;
  unsigned int *sched;
  int skip;
  DvrRenderStats *stats;    // instrumented variant only
  unsigned int *cellBitmap; // instrumented variant only
  SyncDev sync;             // sync variant: signal peers when the partial image is complete
};

struct ResolveLaunch
{
  uint32_t width, height;
  int format, frameID;
  float4 background;
  // background image (0 = constant colour) and what is needed to regenerate the pixel-sample's screen coordinate
  cudaTextureObject_t bgTex;
  float invW, invH;
  int centered;
  BuffersDev fb;
  const float4 *partialRgba;
  const float *partialDepth;
  uint32_t objId, instId;
  size_t pixelBegin, pixelEnd;
};

struct PeerResolveLaunch
{
  ResolveLaunch r;
  CameraDev cam;
  float invW, invH;
  int nSlabs;
  const float4 *rgba[kMaxSlabs];
  const float *depth[kMaxSlabs];
  // sync variant
  SyncDev sync;
  int cull;                   // regenerate the primary ray and skip peer loads when it misses the bounds
  int missValid, missX0, missY0, missX1, missY1; // pixels outside this rectangle miss without regenerating the ray
  int integrator;
  float3 boundsLo, boundsHi;
  float xfm[12];
  uint32_t identity;
};

// Sort-last frame in ONE launch per GPU: march of the own slab, per-region completion flags to the regions' owners,
// compositing + resolve of the owned regions over peer memory, background pixels of the own strip.
struct SlabFrameLaunch
{
  PartialLaunch m;        // march part (own slab -> own partial image)
  PeerResolveLaunch c;    // composite part: all ranks' partial images of this frame, resolve parameters, camera
  uint32_t nRanks, rank;
  uint32_t seq;           // frame number carried by every flag of this launch
  uint32_t tilesPerRegion, nRegions;
  unsigned int *regionDone;                 // local per-region tile counters, zero on entry, re-armed by the last warp
  unsigned int *regionFlags[kMaxSlabs];     // rank p's region table [maxRegions][kMaxSlabs]; we store [region][rank]
  const unsigned int *myRegionFlags;        // the local table: [region][src] >= seq once src finished the region
  unsigned int *resolvedFlags[kMaxSlabs];   // rank p's "rank q resolved frame seq" table [kMaxSlabs]; we store [rank]
  const unsigned int *myResolved;
  int waitAllResolved;                      // display rank: do not retire before every rank's strip has landed
  size_t bgPixelBegin, bgPixelEnd;          // this rank's share of the pixels outside the tile window
  unsigned long long *timing;               // optional %globaltimer stamps (DvrSlabExchange::timing)
  unsigned spinSleepNs, spinSleepCapNs;     // back-off of the warps that wait for region flags: first interval, cap of the doubling
  unsigned debugFlags;                      // timing experiments only (DVR_B200_SLAB_DEBUG; 8: also composite between march tiles, 16: no background chunk between march tiles): 1 no region flags, 2 no
                                            // background strip, 4 no compositing — frames are then incomplete
};

// error plumbing -----------------------------------------------------------------------------
void setError(const std::string &msg);
int cudaFail(cudaError_t e, const char *what);
#define DVR_CUDA(call)                                                                                    \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return ::dvr::cudaFail(_e, #call);                                                                  \
  } while (0)

void countLaunch(unsigned n = 1);
unsigned int *acquireSchedSlot(); // device pair of counters, zero on entry to every launch
int smCount();
// stream-ordered scratch from the library's own pool (released with cudaFreeAsync)
cudaError_t scratchAllocAsync(void **p, size_t bytes, cudaStream_t stream);

// launchers (defined in the .cu files) --------------------------------------------------------
int launchFrame(const FrameLaunch &p, bool skip, bool stats, cudaStream_t s);
int launchBackgroundSweep(const FrameLaunch &p, cudaStream_t s);
// mixed scenes (dvr_scene.cu): a world that also holds surfaces and lights
bool sceneHasSurfaces(const DvrSceneParams *scene);
int launchSceneFrame(const FrameLaunch &p, const DvrSceneParams *scene, bool skip, cudaStream_t s);
int launchPartial(const PartialLaunch &p, cudaStream_t s);
int launchResolve(const ResolveLaunch &p, cudaStream_t s);
int launchCompositeOver(float4 *front, float *frontDepth, const float4 *back, const float *backDepth,
    size_t begin, size_t end, bool backIsInFront, cudaStream_t s);
int launchPeerResolve(const PeerResolveLaunch &p, cudaStream_t s);
int launchSlabFrame(const SlabFrameLaunch &p, cudaStream_t s);
int launchSignalFlags(const SyncDev &sy, cudaStream_t s);
int launchWaitFlags(const unsigned int *flags, uint32_t n, uint32_t value, unsigned int *errorFlag, cudaStream_t s);
int launchScaleVec3(const float *in, float *out, size_t n, float scale, cudaStream_t s);
int launchMacrocellBuild(cudaTextureObject_t pointTex, int3 dims, int zTexBegin, int texDepth, int3 gridDims,
    float2 *ranges, cudaStream_t s);
bool macrocellLinearIsVectorisable(const void *voxels, int3 dims);
int launchMacrocellBuildLinear(const float *voxels, int3 dims, int3 gridDims, float2 *ranges,
    cudaSurfaceObject_t uploadTo, cudaStream_t s);
int launchNvdbBrickBuild(const NvdbDev &g, bool quant, int3 org, int3 dims, int2 *table, float *bricks,
    unsigned int *counter, unsigned int capacity, cudaStream_t s);
int launchMacrocellBuildNvdb(const FieldDev &f, float2 *ranges, cudaStream_t s);
// value ranges on the delta-tracking grid: gridDims cells dividing `spanVoxels` voxel units evenly per axis
int launchDdaRangeBuild(const FieldDev &f, cudaTextureObject_t pointTex, int3 gridDims, float3 cellWidthVoxels,
    float2 *ranges, cudaStream_t s);
int launchCountEmpty(const float *maxOpacities, size_t n, unsigned long long *out, cudaStream_t s);
int launchSelftestLattice(uint32_t count, unsigned long long seed, unsigned int *mismatches, cudaStream_t s);
int launchReferenceGridBuild(const FieldDev &f, int3 gridDims, const float4 *tf, float2 *ranges, float *maxOpacities,
    cudaStream_t s);
int launchMajorants(const float2 *ranges, size_t nCells, const float4 *tf, float vrLo, float vrHi,
    float *maxOpacities, cudaStream_t s);
int launchMajorantsCoarse(const float *fine, int3 gridDims, float *coarse, int3 coarseDims, cudaStream_t s);
int launchRangeReduce(const float2 *ranges, size_t nCells, float2 *out, cudaStream_t s);
int launchPopcount(const unsigned int *bitmap, size_t nWords, unsigned long long *out, cudaStream_t s);
int launchConvertToFloat(const void *src, int dataType, float *dst, size_t n, cudaStream_t s);

} // namespace dvr
