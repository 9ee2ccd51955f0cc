// dvr_nvdb_bricks.cu — apron bricks for NanoVDB fields (see NvdbDev in dvr_nanovdb.cuh).
//
// The reference samples a NanoVDB grid with eight Tree::getValue calls per lattice point through a ReadAccessor
// (gpu/sampleSpatialField.h:80-109).  The values those calls return depend only on the voxel, so they are
// gathered once per field into a layout a marching thread can address without walking the tree: per 8^3 cell one
// 9^3 float brick (the cell plus the +1 apron its trilinear stencils reach), or a single constant when all 729
// values are bit-identical (tiles, background).  Same floats, same interpolation arithmetic: frames are
// bit-identical to the tree walk (tests/test_gpu_nanovdb.py).
#include "dvr_internal.h"

namespace dvr {

template <bool QUANT>
__global__ void __launch_bounds__(256) dvrNvdbBrickBuildKernel(const __grid_constant__ NvdbDev g, int3 org, int3 dims,
    int2 *__restrict__ table, float *__restrict__ bricks, unsigned int *__restrict__ counter, unsigned int capacity)
{
  __shared__ float vals[kNvdbBrickVoxels];
  __shared__ unsigned int slot;
  const unsigned int cell = blockIdx.x;
  const int tz = (int)(cell % (unsigned)dims.z), ty = (int)((cell / (unsigned)dims.z) % (unsigned)dims.y),
            tx = (int)(cell / ((unsigned)dims.z * (unsigned)dims.y));
  const int x0 = (org.x + tx) * 8, y0 = (org.y + ty) * 8, z0 = (org.z + tz) * 8;
  NvdbCache c;
  c.reset();
  bool same = true;
  for (int n = threadIdx.x; n < kNvdbBrickVoxels; n += blockDim.x) {
    const int lx = n / (kNvdbBrickEdge * kNvdbBrickEdge), ly = (n / kNvdbBrickEdge) % kNvdbBrickEdge,
              lz = n % kNvdbBrickEdge;
    vals[n] = nvdbGetValue<QUANT>(g, c, x0 + lx, y0 + ly, z0 + lz);
  }
  __syncthreads();
  const int first = __float_as_int(vals[0]);
  for (int n = threadIdx.x; n < kNvdbBrickVoxels; n += blockDim.x)
    same &= __float_as_int(vals[n]) == first;
  if (__syncthreads_and(same)) {
    if (threadIdx.x == 0)
      table[cell] = make_int2(-1, first);
    return;
  }
  if (threadIdx.x == 0)
    slot = atomicAdd(counter, 1u);
  __syncthreads();
  if (!bricks || slot >= capacity) // counting pass
    return;
  float *dst = bricks + (size_t)slot * kNvdbBrickVoxels;
  for (int n = threadIdx.x; n < kNvdbBrickVoxels; n += blockDim.x)
    dst[n] = vals[n];
  if (threadIdx.x == 0)
    table[cell] = make_int2((int)slot, 0);
}

// counter: one device word, zeroed by the caller.  bricks == nullptr: only count the cells that need a brick.
int launchNvdbBrickBuild(const NvdbDev &g, bool quant, int3 org, int3 dims, int2 *table, float *bricks,
    unsigned int *counter, unsigned int capacity, cudaStream_t s)
{
  const unsigned int cells = (unsigned int)dims.x * (unsigned int)dims.y * (unsigned int)dims.z;
  if (quant)
    dvrNvdbBrickBuildKernel<true><<<cells, 256, 0, s>>>(g, org, dims, table, bricks, counter, capacity);
  else
    dvrNvdbBrickBuildKernel<false><<<cells, 256, 0, s>>>(g, org, dims, table, bricks, counter, capacity);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

} // namespace dvr
