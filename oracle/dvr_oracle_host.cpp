// dvr_oracle_host.cpp — O-cpu, host-side parameter helpers (camera set-up, TF discretisation).
// TEST INFRASTRUCTURE ONLY (see dvr_oracle.cpp).  Compiled with -ffp-contract=off: the reference's
// host code is built for baseline x86-64 (no FMA), so these must not contract either; the
// device-emulating half (dvr_oracle.cpp) is compiled WITH contraction like nvcc's default.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "dvr_oracle.h"

namespace {
struct V3
{
  float x, y, z;
};
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 v3(const float *p) { return {p[0], p[1], p[2]}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normalize(V3 v)
{
  const float d = v.x * v.x + v.y * v.y + v.z * v.z;
  return v * (1.0f / std::sqrt(d));
}
inline float position(float v, float lo, float hi)
{ // gpu_math.h:176-180
  v = std::fmax(lo, std::fmin(v, hi));
  return (v - lo) * (1.f / (hi - lo));
}
} // namespace

extern "C" {

int oracle_camera_perspective(const float pos[3], const float dir[3], const float up[3], float fovy, float aspect,
    float focusDistance, float apertureRadius, const float region[4], DvrCamera *out)
{ // camera/Perspective.cpp:42-72, camera/Camera.cpp:68-76
  std::memset(out, 0, sizeof(*out));
  const float reg[4] = {region ? region[0] : 0.f, region ? region[1] : 0.f, region ? region[2] : 1.f,
      region ? region[3] : 1.f};
  std::memcpy(out->region, reg, sizeof(reg));
  const V3 d = normalize(v3(dir)), u = normalize(v3(up));
  out->type = DVR_CAMERA_PERSPECTIVE;
  const float sy = 2.f * std::tan(0.5f * fovy), sx = sy * aspect;
  V3 du = normalize(cross(d, u)) * sx;
  V3 dv = normalize(cross(du, d)) * sy;
  V3 d00 = d - .5f * du - .5f * dv;
  const float ap = apertureRadius / (sx * focusDistance);
  if (ap > 0.f) {
    du = du * focusDistance;
    dv = dv * focusDistance;
    d00 = d00 * focusDistance;
  }
  auto st = [](float *o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
  st(out->pos, v3(pos));
  st(out->dir, d);
  st(out->up, u);
  st(out->du, du);
  st(out->dv, dv);
  st(out->p00, d00);
  out->scaledAperture = ap;
  out->aspect = aspect;
  return 0;
}

int oracle_camera_orthographic(const float pos[3], const float dir[3], const float up[3], float height, float aspect,
    const float region[4], DvrCamera *out)
{ // camera/Orthographic.cpp:38-52
  std::memset(out, 0, sizeof(*out));
  const float reg[4] = {region ? region[0] : 0.f, region ? region[1] : 0.f, region ? region[2] : 1.f,
      region ? region[3] : 1.f};
  std::memcpy(out->region, reg, sizeof(reg));
  const V3 d = normalize(v3(dir)), u = normalize(v3(up)), p = v3(pos);
  out->type = DVR_CAMERA_ORTHOGRAPHIC;
  const V3 du = normalize(cross(d, u)) * (height * aspect);
  const V3 dv = normalize(cross(du, d)) * height;
  const V3 p00 = p - 0.5f * du - 0.5f * dv;
  auto st = [](float *o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
  st(out->pos, p);
  st(out->dir, d);
  st(out->up, u);
  st(out->du, du);
  st(out->dv, dv);
  st(out->p00, p00);
  out->aspect = aspect;
  return 0;
}

// TransferFunction1D::discritizeTFData, TransferFunction1D.cpp:101-150
int oracle_tf_discretize(const float *color, size_t nColor, int colorChannels, const float *opacity,
    size_t nOpacity, const float uniformColor[4], float uniformOpacity, const float valueRange[2], float *outRgba)
{
  const float lo = valueRange[0], hi = valueRange[1];
  auto positions = [&](size_t n) { // colorMapHelpers.h:43-59
    std::vector<float> p(n);
    p.front() = 0.f;
    p.back() = 1.f;
    const float w = 1.f / (n - 1);
    for (int i = 1; i < (int)n - 1; i++)
      p[i] = p[i - 1] + w;
    for (auto &v : p)
      v = v * (hi - lo) + lo;
    return p;
  };
  auto interp = [&](const float *vals, int nch, const std::vector<float> &pos, float x, float *out) {
    for (size_t i = 0; i + 1 < pos.size(); i++) { // colorMapHelpers.h:61-72
      const float r0 = position(pos[i], lo, hi), r1 = position(pos[i + 1], lo, hi);
      if (x >= r0 && x <= r1) {
        const float a = position(x, r0, r1);
        for (int c = 0; c < nch; ++c)
          out[c] = vals[i * nch + c] * (1.f - a) + vals[(i + 1) * nch + c] * a;
        return;
      }
    }
    const float *src = x <= position(pos[0], lo, hi) ? vals : vals + (pos.size() - 1) * nch;
    for (int c = 0; c < nch; ++c)
      out[c] = src[c];
  };
  std::vector<float> cp, op;
  if (color)
    cp = positions(nColor);
  if (opacity)
    op = positions(nOpacity);
  for (size_t i = 0; i < DVR_TF_SIZE; ++i) {
    const float p = float(i) / (DVR_TF_SIZE - 1);
    float c[4] = {uniformColor[0], uniformColor[1], uniformColor[2], uniformColor[3]};
    if (color) {
      interp(color, colorChannels, cp, p, c);
      if (colorChannels == 3)
        c[3] = 1.f;
    }
    float o = uniformOpacity;
    if (opacity)
      interp(opacity, 1, op, p, &o);
    outRgba[4 * i + 0] = c[0];
    outRgba[4 * i + 1] = c[1];
    outRgba[4 * i + 2] = c[2];
    outRgba[4 * i + 3] = c[3] * o;
  }
  return 0;
}

} // extern "C"
