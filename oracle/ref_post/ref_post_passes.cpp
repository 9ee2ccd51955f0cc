// ref_post_passes.cpp — TEST INFRASTRUCTURE.  The reference's frame post passes compiled IN PLACE (serial
// `parallel_for` fallback of tsd/src/render_pipeline/passes/detail/parallel_for.h: no ENABLE_CUDA / ENABLE_TBB), behind
// a C entry point per pass.  Nothing is copied: the three sources are included where they lie under $(REF).
//   OutlineRenderPass.cpp   computeOutline + shadePixel      -> refpost_outline
//   VisualizeDepthPass.cpp  computeDepthImage                -> refpost_visualize_depth
//   RenderPass.cpp          base class (dimensions, friend RenderPipeline)
// Only tests/ load the resulting oracle/_ref/libref_post.so (the checker of csrc/dvr_post.cu and oracle/post_oracle.py).
#include "render_pipeline/passes/RenderPass.cpp"
#include "render_pipeline/passes/OutlineRenderPass.cpp"
#include "render_pipeline/passes/VisualizeDepthPass.cpp"

namespace tsd {
// RenderPass befriends the pipeline that drives it (RenderPass.h): the same two calls RenderPipeline::render makes
struct RenderPipeline
{
  static void run(RenderPass &pass, RenderPass::Buffers &b, uint32_t w, uint32_t h, int stageId)
  {
    pass.setDimensions(w, h);
    pass.render(b, stageId);
  }
};
} // namespace tsd

extern "C" {

void refpost_outline(uint32_t *color, const uint32_t *objectId, uint32_t w, uint32_t h, uint32_t outlineId)
{
  tsd::OutlineRenderPass pass;
  pass.setOutlineId(outlineId);
  tsd::RenderPass::Buffers b;
  b.color = color;
  b.objectId = const_cast<uint32_t *>(objectId);
  tsd::RenderPipeline::run(pass, b, w, h, 1);
}

void refpost_visualize_depth(uint32_t *color, const float *depth, uint32_t w, uint32_t h, float maxDepth)
{
  tsd::VisualizeDepthPass pass;
  pass.setMaxDepth(maxDepth);
  tsd::RenderPass::Buffers b;
  b.color = color;
  b.depth = const_cast<float *>(depth);
  tsd::RenderPipeline::run(pass, b, w, h, 1);
}

} // extern "C"
