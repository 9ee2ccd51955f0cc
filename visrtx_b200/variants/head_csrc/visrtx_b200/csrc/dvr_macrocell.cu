// dvr_macrocell.cu — K4 (macrocell value ranges), K5 (per-macrocell majorants from the transfer
// function) and small reductions.  Replaces space_skipping/UniformGrid.cu:44-258 of the reference.
//
// K4 is one pass over the voxels (read through a point-sampled view of the same 3-D array the
// marcher filters, so fixed-point formats are seen exactly as the filter sees them).  Cell c
// covers voxel indices [16c-1, 16c+17] per axis (clamped): every trilinear fetch whose lower tap
// index falls in [16c, 16c+15] reads voxels [16c, 16c+16]; the extra voxel on either side absorbs
// the difference between the marcher's fp32 cell test and the texture unit's own coordinate rounding.  The reference's build samples the wrong coordinates
//  and is deliberately not reproduced.
#include <algorithm>

#include "dvr_internal.h"
#include "dvr_march.cuh"

namespace dvr {

__global__ void __launch_bounds__(256) dvrMacrocellRangeKernel(cudaTextureObject_t pointTex, int3 dims,
    int zTexBegin, int texDepth, int3 gridDims, float2 *__restrict__ ranges)
{
  const int cx = blockIdx.x, cy = blockIdx.y, cz = blockIdx.z;
  const int x0 = max(cx * 16 - 1, 0), x1 = min(cx * 16 + 17, dims.x - 1);
  const int y0 = max(cy * 16 - 1, 0), y1 = min(cy * 16 + 17, dims.y - 1);
  const int z0 = max(cz * 16 - 1, 0), z1 = min(cz * 16 + 17, dims.z - 1);
  const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
  const int n = nx * ny * nz;
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = x0 + i % nx, y = y0 + (i / nx) % ny, z = z0 + i / (nx * ny);
    // slices outside the resident slab are clamped by the texture: conservative only for the
    // slices this rank samples, which is all it is used for
    const int zl = min(max(z - zTexBegin, 0), texDepth - 1);
    const float v = tex3D<float>(pointTex, (float)x + 0.5f, (float)y + 0.5f, (float)zl + 0.5f);
    lo = fminf(lo, v); // fminf/fmaxf drop NaNs
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[8], shi[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    slo[w] = lo;
    shi[w] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      lo = fminf(lo, slo[i]);
      hi = fmaxf(hi, shi[i]);
    }
    ranges[((size_t)cz * gridDims.y + cy) * gridDims.x + cx] = make_float2(lo, hi);
  }
}

int launchMacrocellBuild(cudaTextureObject_t pointTex, int3 dims, int zTexBegin, int texDepth, int3 gridDims,
    float2 *ranges, cudaStream_t s)
{
  dim3 grid(gridDims.x, gridDims.y, gridDims.z);
  dvrMacrocellRangeKernel<<<grid, 256, 0, s>>>(pointTex, dims, zTexBegin, texDepth, gridDims, ranges);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// ---- K4, separable variant for f32 fields that arrive as LINEAR device memory (in-situ / time-varying fields) ---
// The apron window [16c-1, 16c+17] is a box, so min/max separate: x (reads every voxel once, coalesced), then y,
// then z over ever smaller intermediates.  Same values as the texture variant (min/max of the same floats), a
// fraction of its time: the field re-finalisation of a time-varying volume is dominated by this build.
__global__ void __launch_bounds__(256) dvrRangeXKernel(const float *__restrict__ vox, int3 dims, int gx, int z0, int nz,
    float2 *__restrict__ outX)
{ // one thread per (cx, y, z): consecutive threads = consecutive cells of one row
  const size_t n = (size_t)gx * dims.y * nz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int cx = (int)(i % gx), y = (int)((i / gx) % dims.y), z = z0 + (int)(i / ((size_t)gx * dims.y));
    const int x0 = max(cx * 16 - 1, 0), x1 = min(cx * 16 + 17, dims.x - 1);
    const float *row = vox + ((size_t)z * dims.y + y) * dims.x;
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int x = x0; x <= x1; ++x) {
      const float v = __ldg(row + x);
      lo = fminf(lo, v); // fminf / fmaxf drop NaNs like the texture variant
      hi = fmaxf(hi, v);
    }
    outX[i] = make_float2(lo, hi);
  }
}

__global__ void __launch_bounds__(256) dvrRangeYKernel(const float2 *__restrict__ inX, int ny, int gx, int gy, int nz,
    float2 *__restrict__ outY)
{ // one thread per (cx, cy, z)
  const size_t n = (size_t)gx * gy * nz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int cx = (int)(i % gx), cy = (int)((i / gx) % gy), z = (int)(i / ((size_t)gx * gy));
    const int y0 = max(cy * 16 - 1, 0), y1 = min(cy * 16 + 17, ny - 1);
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int y = y0; y <= y1; ++y) {
      const float2 r = inX[((size_t)z * ny + y) * gx + cx];
      lo = fminf(lo, r.x);
      hi = fmaxf(hi, r.y);
    }
    outY[i] = make_float2(lo, hi);
  }
}

__global__ void __launch_bounds__(256) dvrRangeZKernel(const float2 *__restrict__ inY, int nz, int3 g, float2 *__restrict__ ranges)
{ // one thread per cell
  const size_t n = (size_t)g.x * g.y * g.z;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int cz = (int)(i / ((size_t)g.x * g.y));
    const size_t xy = i % ((size_t)g.x * g.y);
    const int z0 = max(cz * 16 - 1, 0), z1 = min(cz * 16 + 17, nz - 1);
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int z = z0; z <= z1; ++z) {
      const float2 r = inY[(size_t)z * g.x * g.y + xy];
      lo = fminf(lo, r.x);
      hi = fmaxf(hi, r.y);
    }
    ranges[i] = make_float2(lo, hi);
  }
}

// x and y passes fused, optionally with the upload into the 3-D array (UPLOAD): one CTA per (128-voxel segment,
// 16-row band, kRangeSlicesPerCta slices) = 8 cells of each of those slices.  Its 8 warps share the band's <= 19 rows (apron rows included); every
// lane loads one float4 (full 512-byte requests), stores it to the array through the surface when the row belongs to
// the band proper, and the x reduction runs on shuffles (aprons: voxel 16c-1 from the lane on the left,
// 16c+16 / 16c+17 from the next group, from memory only at segment ends); rows are then folded in registers
// and across warps in shared memory.  The apron rows are fetched by two bands, close in launch order, from L2.
constexpr int kRangeSlicesPerCta = 4;

template <bool UPLOAD>
__global__ void __launch_bounds__(256) dvrRangeXYKernel(const float *__restrict__ vox, cudaSurfaceObject_t surf,
    int3 dims, int gx, int gy, float2 *__restrict__ outY)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nseg = (dims.x + 127) >> 7;
  const int seg = blockIdx.x % nseg;
  const int cy = (blockIdx.x / nseg) % gy;
  const int zFirst = (blockIdx.x / (nseg * gy)) * kRangeSlicesPerCta;
  const int zLast = min(zFirst + kRangeSlicesPerCta, dims.z) - 1;
  const int y0 = max(cy * 16 - 1, 0), y1 = min(cy * 16 + 17, dims.y - 1);
  const int x = (seg << 7) + (lane << 2);
  const bool in = x < dims.x;
  const int nextLane = (lane & ~3) + 4;
  __shared__ float2 part[2][8][8];

  // this warp's rows of one slice: y0 + warp + 8k, k < 3.  The loads of slice z+1 are in flight while slice z is
  // stored and reduced (the pass is latency-bound otherwise: one dependent load -> store chain per warp).
  float4 cur[3], nxt[3];
  auto fetch = [&](int z, float4 (&v)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int y = y0 + warp + 8 * k;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in && y <= y1)
        v[k] = __ldg(reinterpret_cast<const float4 *>(vox + ((size_t)z * dims.y + y) * dims.x + x));
    }
  };
  fetch(zFirst, cur);
  for (int z = zFirst; z <= zLast; ++z) {
    if (z < zLast)
      fetch(z + 1, nxt);
    float accLo = FLT_MAX, accHi = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int y = y0 + warp + 8 * k;
      if (y > y1) // warp-uniform
        break;
      const float4 v = cur[k];
      const float *row = vox + ((size_t)z * dims.y + y) * dims.x;
      if (UPLOAD && in && (y >> 4) == cy)
        surf3Dwrite(v, surf, x * (int)sizeof(float), y, z);
      float lo = in ? fminf(fminf(v.x, v.y), fminf(v.z, v.w)) : FLT_MAX;
      float hi = in ? fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) : -FLT_MAX;
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 1));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, 2));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
      const float left = __shfl_up_sync(0xffffffffu, v.w, 1);
      const float r0 = __shfl_sync(0xffffffffu, v.x, nextLane & 31);
      const float r1 = __shfl_sync(0xffffffffu, v.y, nextLane & 31);
      if ((lane & 3) == 0 && in) {
        if (lane > 0) {
          lo = fminf(lo, left);
          hi = fmaxf(hi, left);
        } else if (x > 0) {
          const float l = __ldg(row + x - 1);
          lo = fminf(lo, l);
          hi = fmaxf(hi, l);
        }
        if (x + 16 < dims.x) { // dims.x % 4 == 0: 16c+16 in range implies 16c+17 in range
          float a = r0, b = r1;
          if (nextLane == 32) {
            a = __ldg(row + x + 16);
            b = __ldg(row + x + 17);
          }
          lo = fminf(fminf(lo, a), b);
          hi = fmaxf(fmaxf(hi, a), b);
        }
        accLo = fminf(accLo, lo);
        accHi = fmaxf(accHi, hi);
      }
    }
    float2(*p)[8] = part[z & 1]; // double-buffered: one barrier per slice
    if ((lane & 3) == 0)
      p[warp][lane >> 2] = make_float2(accLo, accHi);
    __syncthreads();
    if (threadIdx.x < 8) {
      const int cx = seg * 8 + threadIdx.x;
      if (cx < gx) {
        float lo = FLT_MAX, hi = -FLT_MAX;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          lo = fminf(lo, p[w][threadIdx.x].x);
          hi = fmaxf(hi, p[w][threadIdx.x].y);
        }
        outY[((size_t)z * gy + cy) * gx + cx] = make_float2(lo, hi);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      cur[k] = nxt[k];
  }
}

bool macrocellLinearIsVectorisable(const void *voxels, int3 dims)
{
  return dims.x % 4 == 0 && (reinterpret_cast<uintptr_t>(voxels) & 15u) == 0
      && (size_t)((dims.x + 127) / 128) * ((dims.y + 15) / 16) * dims.z < 0x7fffffffull; // grid.x
}

// voxels: linear f32 device memory, x fastest, the WHOLE field (no slab).  uploadTo != 0 (only when
// macrocellLinearIsVectorisable): the same pass also writes the voxels into that surface (the field's 3-D array), so
// a field refresh reads its input once.  Scratch: gx*gy*nz float2 for the y-reduced slices (+ one z-chunk of
// x-reduced rows on the scalar route).
int launchMacrocellBuildLinear(const float *voxels, int3 dims, int3 gridDims, float2 *ranges,
    cudaSurfaceObject_t uploadTo, cudaStream_t s)
{
  const int gx = gridDims.x, gy = gridDims.y;
  const bool vec = macrocellLinearIsVectorisable(voxels, dims);
  if (uploadTo && !vec)
    return cudaFail(cudaErrorInvalidValue, "macrocell build (linear): fused upload needs the vectorised route");
  const int zChunk = 32;
  float2 *bufX = nullptr, *bufY = nullptr;
  if (!vec)
    DVR_CUDA(scratchAllocAsync((void **)&bufX, (size_t)gx * dims.y * zChunk * sizeof(float2), s));
  cudaError_t e = scratchAllocAsync((void **)&bufY, (size_t)gx * gy * dims.z * sizeof(float2), s);
  if (e != cudaSuccess) {
    if (bufX)
      cudaFreeAsync(bufX, s);
    return cudaFail(e, "cudaMallocAsync(macrocell scratch)");
  }
  const unsigned cap = (unsigned)smCount() * 16u;
  if (vec) {
    const unsigned units = (unsigned)((size_t)((dims.x + 127) / 128) * gy
        * ((dims.z + kRangeSlicesPerCta - 1) / kRangeSlicesPerCta));
    if (uploadTo)
      dvrRangeXYKernel<true><<<units, 256, 0, s>>>(voxels, uploadTo, dims, gx, gy, bufY);
    else
      dvrRangeXYKernel<false><<<units, 256, 0, s>>>(voxels, 0, dims, gx, gy, bufY);
    countLaunch();
  } else {
    for (int z0 = 0; z0 < dims.z; z0 += zChunk) {
      const int nz = min(zChunk, dims.z - z0);
      const size_t nX = (size_t)gx * dims.y * nz, nY = (size_t)gx * gy * nz;
      dvrRangeXKernel<<<(unsigned)std::min<size_t>((nX + 255) / 256, cap), 256, 0, s>>>(voxels, dims, gx, z0, nz, bufX);
      dvrRangeYKernel<<<(unsigned)std::min<size_t>((nY + 255) / 256, cap), 256, 0, s>>>(bufX, dims.y, gx, gy, nz,
          bufY + (size_t)z0 * gx * gy);
      countLaunch(2);
    }
  }
  const size_t nC = (size_t)gx * gy * gridDims.z;
  dvrRangeZKernel<<<(unsigned)std::min<size_t>((nC + 255) / 256, cap), 256, 0, s>>>(bufY, dims.z, gridDims, ranges);
  countLaunch();
  e = cudaGetLastError();
  if (bufX)
    cudaFreeAsync(bufX, s);
  cudaFreeAsync(bufY, s);
  if (e != cudaSuccess)
    return cudaFail(e, "macrocell build (linear)");
  return DVR_OK;
}

// K4 for NanoVDB grids: same cell definition in index space relative to the index bounding box; voxels
// outside the box are read through the tree (tile / background values), exactly what a fetch would see.
__global__ void __launch_bounds__(256) dvrMacrocellRangeNvdbKernel(const __grid_constant__ FieldDev f,
    float2 *__restrict__ ranges)
{
  const int cx = blockIdx.x, cy = blockIdx.y, cz = blockIdx.z;
  const int x0 = cx * 16 - 1, y0 = cy * 16 - 1, z0 = cz * 16 - 1;
  const int nn = 19;
  NvdbCache cache;
  cache.reset();
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int i = threadIdx.x; i < nn * nn * nn; i += blockDim.x) {
    // z fastest (NanoVDB leaf order) so consecutive threads share leaves
    const int z = z0 + i % nn, y = y0 + (i / nn) % nn, x = x0 + i / (nn * nn);
    const float v = f.kind == FIELD_NANOVDB_QUANT
        ? nvdbGetValue<true>(f.nv, cache, x + f.nv.bboxMin.x, y + f.nv.bboxMin.y, z + f.nv.bboxMin.z)
        : nvdbGetValue<false>(f.nv, cache, x + f.nv.bboxMin.x, y + f.nv.bboxMin.y, z + f.nv.bboxMin.z);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[8], shi[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    slo[w] = lo;
    shi[w] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      lo = fminf(lo, slo[i]);
      hi = fmaxf(hi, shi[i]);
    }
    ranges[((size_t)cz * f.gridDims.y + cy) * f.gridDims.x + cx] = make_float2(lo, hi);
  }
}

int launchMacrocellBuildNvdb(const FieldDev &f, float2 *ranges, cudaStream_t s)
{
  dim3 grid(f.gridDims.x, f.gridDims.y, f.gridDims.z);
  dvrMacrocellRangeNvdbKernel<<<grid, 256, 0, s>>>(f, ranges);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// Value ranges on the delta-tracking (DDA) grid: the reference's geometry — gridDims cells dividing the
// field bounds evenly (UniformGrid.cu:152-154, dda.h) — so a cell is `w` voxel units wide with w generally
// not an integer.  Cell c covers lower-tap indices floor(c*w) .. ceil((c+1)*w); one extra voxel each side.
__global__ void __launch_bounds__(256) dvrDdaRangeKernel(const __grid_constant__ FieldDev f,
    cudaTextureObject_t pointTex, int3 g, float3 w, float2 *__restrict__ ranges)
{
  const int cx = blockIdx.x, cy = blockIdx.y, cz = blockIdx.z;
  int x0 = (int)floorf(cx * w.x) - 1, x1 = (int)ceilf((cx + 1) * w.x) + 1;
  int y0 = (int)floorf(cy * w.y) - 1, y1 = (int)ceilf((cy + 1) * w.y) + 1;
  int z0 = (int)floorf(cz * w.z) - 1, z1 = (int)ceilf((cz + 1) * w.z) + 1;
  const bool nvdb = f.kind >= FIELD_NANOVDB;
  if (!nvdb) {
    x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
    x1 = min(x1, f.dims.x - 1); y1 = min(y1, f.dims.y - 1); z1 = min(z1, f.dims.z - 1);
  }
  const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
  const int n = nx * ny * nz;
  NvdbCache cache;
  cache.reset();
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v;
    if (nvdb) {
      const int z = z0 + i % nz, y = y0 + (i / nz) % ny, x = x0 + i / (nz * ny);
      v = f.kind == FIELD_NANOVDB_QUANT
          ? nvdbGetValue<true>(f.nv, cache, x + f.nv.bboxMin.x, y + f.nv.bboxMin.y, z + f.nv.bboxMin.z)
          : nvdbGetValue<false>(f.nv, cache, x + f.nv.bboxMin.x, y + f.nv.bboxMin.y, z + f.nv.bboxMin.z);
    } else {
      const int x = x0 + i % nx, y = y0 + (i / nx) % ny, z = z0 + i / (nx * ny);
      const int zl = min(max(z - f.zTexBegin, 0), f.texDepth - 1);
      v = tex3D<float>(pointTex, (float)x + 0.5f, (float)y + 0.5f, (float)zl + 0.5f);
    }
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[8], shi[8];
  const int wp = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    slo[wp] = lo;
    shi[wp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      lo = fminf(lo, slo[i]);
      hi = fmaxf(hi, shi[i]);
    }
    ranges[((size_t)cz * g.y + cy) * g.x + cx] = make_float2(lo, hi);
  }
}

int launchDdaRangeBuild(const FieldDev &f, cudaTextureObject_t pointTex, int3 gridDims, float3 cellWidthVoxels,
    float2 *ranges, cudaStream_t s)
{
  dim3 grid(gridDims.x, gridDims.y, gridDims.z);
  dvrDdaRangeKernel<<<grid, 256, 0, s>>>(f, pointTex, gridDims, cellWidthVoxels, ranges);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// ---- the reference's own delta-tracking grid, defects included (opt-in: DvrFrameParams::dptReferenceGrid) -------
// buildGridGPU (UniformGrid.cu:92-150) runs one thread per MACROCELL: it takes the max of the field at the eight
// points (cellID +- .5) / gridDims — coordinates in [0,1] that it hands to the sampler as object-space positions
//  — and splats that one value as both range ends into the cells its bounds project onto.
// computeMaxOpacitiesGPU (UniformGrid.cu:55-90) then classifies the ranges with the DEFAULT value range {0,1}
// (Q8) over texels int(lo*255) .. int(hi*255)+1.  Reproduced operation for operation so that a dpt frame can be
// made to match the real reference bit for bit; the default grid (above) is the conservative one.
__device__ __forceinline__ void atomicMinFloat(float *address, float val)
{ // gpu_util.h:108-117
  int ret = __float_as_int(*address);
  while (val < __int_as_float(ret)) {
    const int old = ret;
    if ((ret = atomicCAS((int *)address, old, __float_as_int(val))) == old)
      break;
  }
}
__device__ __forceinline__ void atomicMaxFloat(float *address, float val)
{ // gpu_util.h:119-128
  int ret = __float_as_int(*address);
  while (val > __int_as_float(ret)) {
    const int old = ret;
    if ((ret = atomicCAS((int *)address, old, __float_as_int(val))) == old)
      break;
  }
}

__global__ void dvrRefGridInvalidateKernel(float2 *__restrict__ ranges, size_t n)
{ // invalidateRangesGPU, UniformGrid.cu:44-53
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n)
    ranges[i] = make_float2(+1e30f, -1e30f);
}

__device__ __forceinline__ int3 projectOnGridRef(float3 V, int3 dims, float3 lo, float3 hi)
{ // uniformGrid.h:44-50
  const float3 v01 = make_float3(__fdiv_rn(__fsub_rn(V.x, lo.x), __fsub_rn(hi.x, lo.x)),
      __fdiv_rn(__fsub_rn(V.y, lo.y), __fsub_rn(hi.y, lo.y)), __fdiv_rn(__fsub_rn(V.z, lo.z), __fsub_rn(hi.z, lo.z)));
  return make_int3(min(max((int)__fmul_rn(v01.x, (float)dims.x), 0), dims.x - 1),
      min(max((int)__fmul_rn(v01.y, (float)dims.y), 0), dims.y - 1),
      min(max((int)__fmul_rn(v01.z, (float)dims.z), 0), dims.z - 1));
}

__global__ void dvrRefGridBuildKernel(const __grid_constant__ FieldDev f, int3 dims, float2 *__restrict__ ranges)
{
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t n = (size_t)dims.x * dims.y * dims.z;
  if (tid >= n)
    return;
  const int3 id = make_int3((int)(tid % dims.x), (int)(tid / dims.x % dims.y), (int)(tid / ((size_t)dims.x * dims.y)));
  const float3 lo = f.boundsLo, hi = f.boundsHi;
  const float3 ext = make_float3(__fdiv_rn(__fsub_rn(hi.x, lo.x), (float)dims.x), __fdiv_rn(__fsub_rn(hi.y, lo.y), (float)dims.y),
      __fdiv_rn(__fsub_rn(hi.z, lo.z), (float)dims.z));
  // voxelBounds: lower + id * ext (one fused multiply-add on the GPU), upper = that + ext
  const float3 bl = make_float3(__fmaf_rn((float)id.x, ext.x, lo.x), __fmaf_rn((float)id.y, ext.y, lo.y),
      __fmaf_rn((float)id.z, ext.z, lo.z));
  const float3 bu = make_float3(__fadd_rn(bl.x, ext.x), __fadd_rn(bl.y, ext.y), __fadd_rn(bl.z, ext.z));
  const float3 halfSpacing = 0.5f * f.spacing;
  NvdbCache cache;
  cache.reset();
  float v = -1e30f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // the corner order of UniformGrid.cu:119-126 does not matter for a max
    const float3 tc = make_float3(__fdiv_rn(__fadd_rn((float)id.x, (k & 1) ? .5f : -.5f), (float)dims.x),
        __fdiv_rn(__fadd_rn((float)id.y, (k & 2) ? .5f : -.5f), (float)dims.y),
        __fdiv_rn(__fadd_rn((float)id.z, (k & 4) ? .5f : -.5f), (float)dims.z));
    float s;
    if (f.kind == FIELD_NANOVDB_QUANT)
      s = nvdbSampleTrilinear<true>(f.nv, cache, nvdbWorldToIndex(f.nv, tc));
    else if (f.kind == FIELD_NANOVDB)
      s = nvdbSampleTrilinear<false>(f.nv, cache, nvdbWorldToIndex(f.nv, tc));
    else {
      const float3 c = fieldTexCoord(f, halfSpacing, tc);
      s = tex3D<float>(f.tex, c.x, c.y, c.z);
    }
    v = fmaxf(v, s);
  }
  const int3 a = projectOnGridRef(bl, dims, lo, hi), b = projectOnGridRef(bu, dims, lo, hi);
  for (int z = a.z; z <= b.z; ++z)
    for (int y = a.y; y <= b.y; ++y)
      for (int x = a.x; x <= b.x; ++x) {
        float2 *r = &ranges[(size_t)z * dims.x * dims.y + (size_t)y * dims.x + x];
        atomicMinFloat(&r->x, v);
        atomicMaxFloat(&r->y, v);
      }
}

__global__ void dvrRefGridMajorantKernel(const float2 *__restrict__ ranges, size_t nCells, const float4 *__restrict__ tf,
    float *__restrict__ maxOpacities)
{ // computeMaxOpacitiesGPU with xfRange = {0,1}; tex1D at texel centres (i + .5)/256 returns the table entry
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c >= nCells)
    return;
  const float2 r = ranges[c];
  if (r.y < r.x) {
    maxOpacities[c] = 0.f;
    return;
  }
  const int lo = min(max((int)__fmul_rn(r.x, 255.f), 0), 255);
  const int hi = min(max((int)__fmul_rn(r.y, 255.f) + 1, 0), 255);
  float m = 0.f;
  for (int i = lo; i <= hi; ++i)
    m = fmaxf(m, __ldg(&tf[i]).w);
  maxOpacities[c] = m;
}

int launchReferenceGridBuild(const FieldDev &f, int3 gridDims, const float4 *tf, float2 *ranges, float *maxOpacities,
    cudaStream_t s)
{
  const size_t n = (size_t)gridDims.x * gridDims.y * gridDims.z;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  dvrRefGridInvalidateKernel<<<blocks, 256, 0, s>>>(ranges, n);
  dvrRefGridBuildKernel<<<blocks, 256, 0, s>>>(f, gridDims, ranges);
  dvrRefGridMajorantKernel<<<blocks, 256, 0, s>>>(ranges, n, tf, maxOpacities);
  DVR_CUDA(cudaGetLastError());
  countLaunch(3);
  return DVR_OK;
}

// K5: majorant = max TF alpha the marcher's lookup can return for any value in the cell's range.
// UniformGrid.cu:55-90 restated with (a) the volume's own valueRange (quirk Q8) and (b) the
// texel interval derived from the same coordinate mapping tfLookup() uses, widened by one texel.
__global__ void dvrMajorantKernel(const float2 *__restrict__ ranges, size_t nCells, const float4 *__restrict__ tf,
    float vrLo, float vrHi, float *__restrict__ maxOpacities)
{
  __shared__ float s_alpha[DVR_TF_SIZE];
  for (int i = threadIdx.x; i < DVR_TF_SIZE; i += blockDim.x)
    s_alpha[i] = tf[i].w;
  __syncthreads();
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c >= nCells)
    return;
  const float2 r = ranges[c];
  if (!(r.x <= r.y)) { // empty / all-NaN cell
    maxOpacities[c] = 0.f;
    return;
  }
  const float c0 = rangePosition(r.x, vrLo, vrHi), c1 = rangePosition(r.y, vrLo, vrHi);
  int i0 = (int)floorf(c0 * 256.0f - 0.5f) - 1;
  int i1 = (int)floorf(c1 * 256.0f - 0.5f) + 2;
  i0 = max(0, min(i0, DVR_TF_SIZE - 1));
  i1 = max(0, min(i1, DVR_TF_SIZE - 1));
  float m = 0.f;
  for (int i = i0; i <= i1; ++i)
    m = fmaxf(m, s_alpha[i]);
  maxOpacities[c] = m;
}

int launchMajorants(const float2 *ranges, size_t nCells, const float4 *tf, float vrLo, float vrHi,
    float *maxOpacities, cudaStream_t s)
{
  if (nCells == 0)
    return DVR_OK;
  dvrMajorantKernel<<<(unsigned)((nCells + 255) / 256), 256, 0, s>>>(ranges, nCells, tf, vrLo, vrHi, maxOpacities);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// coarse level of the skipping hierarchy: max over each 4x4x4 block of macrocells
__global__ void dvrMajorantCoarseKernel(const float *__restrict__ fine, int3 g, float *__restrict__ coarse, int3 cg)
{
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t n = (size_t)cg.x * cg.y * cg.z;
  if (c >= n)
    return;
  const int X = (int)(c % cg.x), Y = (int)((c / cg.x) % cg.y), Z = (int)(c / ((size_t)cg.x * cg.y));
  float m = 0.f;
  for (int z = Z * 4; z < min(Z * 4 + 4, g.z); ++z)
    for (int y = Y * 4; y < min(Y * 4 + 4, g.y); ++y)
      for (int x = X * 4; x < min(X * 4 + 4, g.x); ++x)
        m = fmaxf(m, fine[((size_t)z * g.y + y) * g.x + x]);
  coarse[c] = m;
}

int launchMajorantsCoarse(const float *fine, int3 gridDims, float *coarse, int3 coarseDims, cudaStream_t s)
{
  const size_t n = (size_t)coarseDims.x * coarseDims.y * coarseDims.z;
  if (n == 0)
    return DVR_OK;
  dvrMajorantCoarseKernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(fine, gridDims, coarse, coarseDims);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// global (min,max) over the macrocell ranges: one CTA, strided
__global__ void dvrRangeReduceKernel(const float2 *__restrict__ ranges, size_t nCells, float2 *out)
{
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (size_t i = threadIdx.x; i < nCells; i += blockDim.x) {
    const float2 r = ranges[i];
    if (r.x <= r.y) {
      lo = fminf(lo, r.x);
      hi = fmaxf(hi, r.y);
    }
  }
  __shared__ float slo[1024], shi[1024];
  slo[threadIdx.x] = lo;
  shi[threadIdx.x] = hi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      slo[threadIdx.x] = fminf(slo[threadIdx.x], slo[threadIdx.x + o]);
      shi[threadIdx.x] = fmaxf(shi[threadIdx.x], shi[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *out = make_float2(slo[0], shi[0]);
}

int launchRangeReduce(const float2 *ranges, size_t nCells, float2 *out, cudaStream_t s)
{
  dvrRangeReduceKernel<<<1, 1024, 0, s>>>(ranges, nCells, out);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

__global__ void dvrPopcountKernel(const unsigned int *__restrict__ bitmap, size_t nWords, unsigned long long *out)
{
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nWords; i += (size_t)gridDim.x * blockDim.x)
    c += __popc(bitmap[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c)
    atomicAdd(out, c);
}

__global__ void dvrCountEmptyKernel(const float *__restrict__ maxOpacities, size_t n, unsigned long long *out)
{
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c += maxOpacities[i] <= 0.f ? 1ull : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c)
    atomicAdd(out, c);
}

// number of macrocells whose majorant is 0 (what skipping can exploit), counted on the device
int launchCountEmpty(const float *maxOpacities, size_t n, unsigned long long *out, cudaStream_t s)
{
  if (n == 0)
    return DVR_OK;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 1184)
    blocks = 1184;
  dvrCountEmptyKernel<<<blocks, 256, 0, s>>>(maxOpacities, n, out);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

int launchPopcount(const unsigned int *bitmap, size_t nWords, unsigned long long *out, cudaStream_t s)
{
  if (nWords == 0)
    return DVR_OK;
  unsigned blocks = (unsigned)((nWords + 255) / 256);
  if (blocks > 1024)
    blocks = 1024;
  dvrPopcountKernel<<<blocks, 256, 0, s>>>(bitmap, nWords, out);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

// element conversion for formats the 3-D array cannot hold natively (FLOAT64 -> f32)
__global__ void dvrF64ToF32Kernel(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

int launchConvertToFloat(const void *src, int dataType, float *dst, size_t n, cudaStream_t s)
{
  if (dataType != DVR_FLOAT64) {
    setError("launchConvertToFloat: only FLOAT64 needs conversion");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148 * 16)
    blocks = 148 * 16;
  dvrF64ToF32Kernel<<<blocks, 256, 0, s>>>((const double *)src, dst, n);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

} // namespace dvr
