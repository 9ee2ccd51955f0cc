// dvr_post.cu — frame post passes on device buffers: the kernels TSD's render pipeline runs
// on the channels it maps through ANARI_NV_FRAME_BUFFERS_CUDA (the .cpp files of tsd/src/render_pipeline/passes, there as
// thrust::for_each / thrust::transform over lambdas).  Pure streaming kernels: HBM-bound, 4..16 B per pixel.
//
// Colour conversion helpers: helium::cvt_color_to_float4 / cvt_color_to_uint32 come from the ANARI-SDK
// (helium/helium_math.h, >= 0.15, not vendored with the reference): c / 255.f per byte, and
// uint32(255.f * clamp(f, 0, 1)) per component (truncation), packed r | g << 8 | b << 16 | a << 24.
// linalg's lerp(a, b, t) = a * (1 - t) + b * t, evaluated here with separately rounded operations.
#include "dvr_internal.h"

namespace dvr {

namespace {

constexpr int kPostThreads = 256;

__device__ __forceinline__ uint32_t cvtComponent(float f)
{
  return (uint32_t)__fmul_rn(255.f, fminf(fmaxf(f, 0.f), 1.f));
}

__device__ __forceinline__ uint32_t shadePixel(uint32_t c)
{ // OutlineRenderPass.cpp:13-19 (same function in AnariSceneRenderPass.cpp:22-28)
  const float hl[4] = {1.f, 0.5f, 0.f, 1.f};
  uint32_t out = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float in = __fdiv_rn((float)((c >> (8 * k)) & 0xffu), 255.f);
    const float v = __fadd_rn(__fmul_rn(in, __fsub_rn(1.f, 0.8f)), __fmul_rn(hl[k], 0.8f));
    out |= cvtComponent(v) << (8 * k);
  }
  return out;
}

__global__ void __launch_bounds__(kPostThreads) dvrPostConvertFloatColorKernel(const float4 *__restrict__ in,
    uint32_t *__restrict__ out, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    // uint8_t(clamp(v, 0, 1) * 255), AnariSceneRenderPass.cpp:17-19
    const uint32_t r = (uint32_t)(uint8_t)__fmul_rn(fminf(fmaxf(v.x, 0.f), 1.f), 255.f);
    const uint32_t g = (uint32_t)(uint8_t)__fmul_rn(fminf(fmaxf(v.y, 0.f), 1.f), 255.f);
    const uint32_t b = (uint32_t)(uint8_t)__fmul_rn(fminf(fmaxf(v.z, 0.f), 1.f), 255.f);
    const uint32_t a = (uint32_t)(uint8_t)__fmul_rn(fminf(fmaxf(v.w, 0.f), 1.f), 255.f);
    out[i] = r | (g << 8) | (b << 16) | (a << 24);
  }
}

__global__ void __launch_bounds__(kPostThreads) dvrPostCompositeDepthKernel(uint32_t *__restrict__ colorOut,
    float *__restrict__ depthOut, uint32_t *__restrict__ idOut, const uint32_t *__restrict__ colorIn,
    const float *__restrict__ depthIn, const uint32_t *__restrict__ idIn, size_t n, int firstPass)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float current = depthIn[i];
    if (firstPass || current < depthOut[i]) { // AnariSceneRenderPass.cpp:37-44
      depthOut[i] = current;
      colorOut[i] = colorIn[i];
      if (idIn)
        idOut[i] = idIn[i];
    }
  }
}

__global__ void __launch_bounds__(kPostThreads) dvrPostOutlineKernel(uint32_t *__restrict__ color,
    const uint32_t *__restrict__ objectId, uint32_t w, uint32_t h, uint32_t outlineId)
{
  const size_t n = (size_t)w * h;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i % w);
    int cnt = 0;
    // unsigned arithmetic as written in the reference: max(0u, y - 1) wraps to UINT_MAX on row 0, and then the
    // loop body never runs (OutlineRenderPass.cpp:29-31); same for column 0
    for (uint32_t fy = max(0u, y - 1u); fy <= min(h - 1u, y + 1u); fy++) {
      for (uint32_t fx = max(0u, x - 1u); fx <= min(w - 1u, x + 1u); fx++)
        if (objectId[fx + (size_t)w * fy] == outlineId)
          cnt++;
    }
    if (cnt > 1 && cnt < 8)
      color[i] = shadePixel(color[i]);
  }
}

__global__ void __launch_bounds__(kPostThreads) dvrPostVisualizeDepthKernel(uint32_t *__restrict__ color,
    const float *__restrict__ depth, size_t n, float maxDepth)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(__fdiv_rn(depth[i], maxDepth), 0.f), 1.f); // VisualizeDepthPass.cpp:17-19
    const uint32_t c = cvtComponent(v);
    color[i] = c | (c << 8) | (c << 16) | (255u << 24);
  }
}

unsigned gridFor(size_t n)
{
  const size_t want = (n + kPostThreads - 1) / kPostThreads;
  const size_t cap = (size_t)smCount() * 8; // grid-stride: a multiple of the SM count
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

} // namespace

int launchPostConvertFloatColor(const float *in, uint32_t *out, size_t n, cudaStream_t s)
{
  dvrPostConvertFloatColorKernel<<<gridFor(n), kPostThreads, 0, s>>>((const float4 *)in, out, n);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

int launchPostCompositeDepth(uint32_t *colorOut, float *depthOut, uint32_t *idOut, const uint32_t *colorIn,
    const float *depthIn, const uint32_t *idIn, size_t n, int firstPass, cudaStream_t s)
{
  dvrPostCompositeDepthKernel<<<gridFor(n), kPostThreads, 0, s>>>(colorOut, depthOut, idOut, colorIn, depthIn, idIn, n,
      firstPass);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

int launchPostOutline(uint32_t *color, const uint32_t *objectId, uint32_t w, uint32_t h, uint32_t outlineId, cudaStream_t s)
{
  dvrPostOutlineKernel<<<gridFor((size_t)w * h), kPostThreads, 0, s>>>(color, objectId, w, h, outlineId);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

int launchPostVisualizeDepth(uint32_t *color, const float *depth, size_t n, float maxDepth, cudaStream_t s)
{
  dvrPostVisualizeDepthKernel<<<gridFor(n), kPostThreads, 0, s>>>(color, depth, n, maxDepth);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

} // namespace dvr

using namespace dvr;

extern "C" {

#define POST_REQUIRE(cond, name)                                        \
  do {                                                                  \
    if (!(cond)) {                                                      \
      setError(name ": invalid argument");                              \
      return DVR_ERR_INVALID_ARGUMENT;                                  \
    }                                                                   \
    if (dvr_device_count() <= 0) {                                      \
      setError(name ": no CUDA device (this library has no CPU fallback)"); \
      return DVR_ERR_NO_DEVICE;                                         \
    }                                                                   \
  } while (0)

int dvr_post_convert_float_color(const float *rgbaF32, uint32_t *rgba8, size_t nPixels, void *stream)
{
  POST_REQUIRE(rgbaF32 && rgba8, "dvr_post_convert_float_color");
  if (nPixels == 0)
    return DVR_OK;
  return launchPostConvertFloatColor(rgbaF32, rgba8, nPixels, (cudaStream_t)stream);
}

int dvr_post_composite_depth(uint32_t *colorOut, float *depthOut, uint32_t *idOut, const uint32_t *colorIn,
    const float *depthIn, const uint32_t *idIn, size_t nPixels, int firstPass, void *stream)
{
  POST_REQUIRE(colorOut && depthOut && colorIn && depthIn && (!idIn || idOut), "dvr_post_composite_depth");
  if (nPixels == 0)
    return DVR_OK;
  return launchPostCompositeDepth(colorOut, depthOut, idOut, colorIn, depthIn, idIn, nPixels, firstPass, (cudaStream_t)stream);
}

int dvr_post_outline(uint32_t *rgba8, const uint32_t *objectId, uint32_t width, uint32_t height, uint32_t outlineId,
    void *stream)
{
  POST_REQUIRE(rgba8 && objectId, "dvr_post_outline");
  if (width == 0 || height == 0 || outlineId == ~0u) // OutlineRenderPass.cpp:58-59
    return DVR_OK;
  return launchPostOutline(rgba8, objectId, width, height, outlineId, (cudaStream_t)stream);
}

int dvr_post_visualize_depth(uint32_t *rgba8, const float *depth, size_t nPixels, float maxDepth, void *stream)
{
  POST_REQUIRE(rgba8 && depth, "dvr_post_visualize_depth");
  if (nPixels == 0)
    return DVR_OK;
  return launchPostVisualizeDepth(rgba8, depth, nPixels, maxDepth, (cudaStream_t)stream);
}

int dvr_post_pick(const float *depth, const uint32_t *objectId, uint32_t width, uint32_t height, uint32_t x, uint32_t y,
    float *depthOut, uint32_t *idOut, void *stream)
{
  POST_REQUIRE(depth && depthOut && idOut && x < width && y < height, "dvr_post_pick");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t i = (size_t)y * width + x;
  *idOut = ~0u;
  DVR_CUDA(cudaMemcpyAsync(depthOut, depth + i, sizeof(float), cudaMemcpyDeviceToHost, s));
  if (objectId)
    DVR_CUDA(cudaMemcpyAsync(idOut, objectId + i, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  DVR_CUDA(cudaStreamSynchronize(s));
  return DVR_OK;
}

} // extern "C"
