// Shim for "tsd/core/TSDMath.hpp" — TEST INFRASTRUCTURE (oracle/ref_post): lets the reference's render-pipeline pass
// sources (tsd/src/render_pipeline/passes/*.cpp) compile in place without the ANARI-SDK, which is not in this image.
//
// The real header pulls anari/anari_cpp/ext/linalg.h (sgorsten's linalg, v2.2) and helium/helium_math.h; neither is
// vendored with the reference.  Only what the pass kernels touch is provided, restated from the published sources:
//   linalg:  vec<T,M> aggregates with element-wise operators, vec<T,4>(vec<T,3>, T), vec<T,3>(T),
//            lerp(a, b, t) = a*(1-t) + b*t
//   helium:  cvt_color_to_float4(uint32) = byte/255.f per channel (r = low byte),
//            cvt_color_to_uint32(float) = uint32(255.f * clamp(f, 0, 1)), cvt_color_to_uint32(float4) = r|g<<8|b<<16|a<<24
// These two helium helpers are therefore still "parity unpinned" (restated, not compiled from their source); everything
// else the pass kernels do — loop bounds, the unsigned window arithmetic of computeOutline, comparison and clamp
// semantics, conversion order — is the reference's own code.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

namespace tsd {
namespace math {

struct uint2
{
  uint32_t x, y;
};

struct float3
{
  float x, y, z;
  constexpr float3() : x(0), y(0), z(0) {}
  constexpr explicit float3(float s) : x(s), y(s), z(s) {}
  constexpr float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct float4
{
  float x, y, z, w;
  constexpr float4() : x(0), y(0), z(0), w(0) {}
  constexpr float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
  constexpr float4(const float3 &v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};

constexpr float4 operator*(const float4 &a, float b) { return {a.x * b, a.y * b, a.z * b, a.w * b}; }
constexpr float4 operator+(const float4 &a, const float4 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
// linalg.h: lerp(a,b,t) { return a*(1-t) + b*t; }
constexpr float4 lerp(const float4 &a, const float4 &b, float t) { return a * (1 - t) + b * t; }

} // namespace math

using math::float3;
using math::float4;
using math::uint2;

} // namespace tsd

namespace helium {

inline tsd::float4 cvt_color_to_float4(uint32_t rgba)
{
  const float r = float((rgba >> 0) & 0xff) / 255.f;
  const float g = float((rgba >> 8) & 0xff) / 255.f;
  const float b = float((rgba >> 16) & 0xff) / 255.f;
  const float a = float((rgba >> 24) & 0xff) / 255.f;
  return tsd::float4(r, g, b, a);
}

inline uint32_t cvt_color_to_uint32(const float &f)
{
  return static_cast<uint32_t>(255.f * std::clamp(f, 0.f, 1.f));
}

inline uint32_t cvt_color_to_uint32(const tsd::float4 &v)
{
  return (cvt_color_to_uint32(v.x) << 0) | (cvt_color_to_uint32(v.y) << 8) | (cvt_color_to_uint32(v.z) << 16)
      | (cvt_color_to_uint32(v.w) << 24);
}

} // namespace helium
