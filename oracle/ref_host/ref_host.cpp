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
