// ref_post_scene.cpp — TEST INFRASTRUCTURE.  The two buffer kernels of the reference's AnariSceneRenderPass
// (convertFloatColorBuffer, compositeFrame: tsd/src/render_pipeline/passes/AnariSceneRenderPass.cpp:17-47).  The rest of
// that file drives an ANARI device and needs the ANARI-SDK, so the Makefile cuts the kernel section out of the
// reference file at build time into a temporary directory (REFPOST_KERNELS, removed after the compile; the recipe
// checks that the section still holds exactly these functions) and this file includes it inside namespace tsd, where
// it lives in the original.
#include <algorithm>
#include <cstring>
#include <limits>

#include "render_pipeline/passes/RenderPass.h"
#include "render_pipeline/passes/detail/parallel_for.h"
#include "render_pipeline/passes/detail/parallel_transform.h"

namespace tsd {
#include REFPOST_KERNELS
} // namespace tsd

extern "C" {

void refpost_convert_float_color(const float *rgba, uint8_t *out, size_t totalSize)
{
  tsd::convertFloatColorBuffer(rgba, out, totalSize);
}

void refpost_composite(uint32_t *colorOut, float *depthOut, uint32_t *idOut, const uint32_t *colorIn, const float *depthIn,
    const uint32_t *idIn, uint32_t w, uint32_t h, int firstPass)
{
  tsd::RenderPass::Buffers o, i;
  o.color = colorOut;
  o.depth = depthOut;
  o.objectId = idOut;
  i.color = const_cast<uint32_t *>(colorIn);
  i.depth = const_cast<float *>(depthIn);
  i.objectId = const_cast<uint32_t *>(idIn);
  tsd::compositeFrame(o, i, tsd::uint2{w, h}, firstPass != 0);
}

} // extern "C"
