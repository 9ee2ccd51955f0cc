// Stub of the few OptiX device-API names the reference's gpu/*.h headers mention, so that
// devices/rtx/gpu/{gpu_util,intersectRay,volumeIntegration}.h compile UNMODIFIED with plain nvcc
// (SURVEY Appendix B).  TEST INFRASTRUCTURE (oracle) — never part of the product.
//
// optixTrace is forwarded to refgpu_shim_trace(), defined in ref_gpu_dvr.cu, which performs the
// volume-AABB search that the RT-core traversal + __intersection__/__closesthit__ programs
// (scene/Intersectors_ptx.cu:248-274, gpu/populateHit.h:370-390) perform in the real device.
#pragma once
#include <cuda_runtime.h>
typedef unsigned long long OptixTraversableHandle;
#define OptixVisibilityMask(x) (x)
enum
{
  OPTIX_RAY_FLAG_NONE = 0,
  OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1,
  OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT = 8,
  OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES = 16
};
__device__ inline uint3 optixGetLaunchIndex()
{
  return make_uint3(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y, 0);
}
__device__ inline uint3 optixGetLaunchDimensions()
{
  return make_uint3(gridDim.x * blockDim.x, gridDim.y * blockDim.y, 1);
}
__device__ inline unsigned optixGetPayload_0() { return 0; }
__device__ inline unsigned optixGetPayload_1() { return 0; }
__device__ inline unsigned optixGetPayload_2() { return 0; }
__device__ inline unsigned optixGetPayload_3() { return 0; }
__device__ inline unsigned optixGetPayload_4() { return 0; }

__device__ void refgpu_shim_trace(unsigned long long traversable, float3 org, float3 dir, float tmin, float tmax,
    unsigned ssHi, unsigned ssLo, unsigned dataHi, unsigned dataLo, unsigned bvhSelection, unsigned rayFlags);

__device__ inline void optixTrace(OptixTraversableHandle h, float3 org, float3 dir, float tmin, float tmax,
    float /*time*/, unsigned /*mask*/, unsigned flags, unsigned /*sbtOffset*/, unsigned /*sbtStride*/,
    unsigned /*miss*/, unsigned &u0, unsigned &u1, unsigned &u2, unsigned &u3, unsigned &u4)
{
  // the ray flags tell the shim which programs would run: DISABLE_CLOSESTHIT = a shadow ray (any-hit accumulation),
  // CULL_BACK_FACING_TRIANGLES = the renderer's cullTriangleBackfaces on primary rays
  refgpu_shim_trace(h, org, dir, tmin, tmax, u0, u1, u2, u3, u4, flags);
}
