// dvr_oracle.cpp — O-cpu: CPU restatement of VisRTX's DVR path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; nothing under visrtx_b200/ links or calls it.
//
// It follows the reference (NVIDIA/VisRTX v0.13.0; paths relative to the reference checkout,
// devices/rtx/ prefix omitted) in execution order:
//   createScreenSample          gpu/createScreenSample.h:38-66      (+ cuRAND Philox4x32-10,
//                               /usr/local/cuda/include/curand_kernel.h:888-1037)
//   makePrimaryRay/cameraCreateRay   gpu/cameraCreateRay.h:38-81
//   camera set-up               camera/Perspective.cpp:42-72, camera/Orthographic.cpp:38-52
//   intersectVolume             scene/Intersectors_ptx.cu:248-274
//   rayMarchAllVolumes          gpu/volumeIntegration.h:317-350
//   detail::rayMarchVolume      gpu/volumeIntegration.h:105-165
//   _rayMarchVolume             gpu/volumeIntegration.h:64-103
//   SpatialFieldSampler<tex>    gpu/sampleSpatialField.h:54-78   (software texture unit below)
//   classifySample              gpu/volumeIntegration.h:46-62, gpu/gpu_math.h:176-180
//   TF discretisation           scene/volume/TransferFunction1D.cpp:101-150, utility/colorMapHelpers.h:43-72
//   raygen volume branch        renderer/Raycast_ptx.cu:139-178, renderer/DirectLight_ptx.cu:376-417
//   accumResults                gpu/gpu_util.h:321-443
//
// PARITY PIN: the reference's own tests hold no golden vector for this path, so the
// pin is (1) O-gpu — the reference's unmodified device headers compiled for sm_100a
// (oracle/ref_gpu, built into oracle/_ref/) — rendered on a B200 and committed as fixtures under
// tests/golden/, which tests/test_oracle_golden.py checks this file against, and (2) the texture
// unit model measured on B200 (profiles/texture_unit_model.md).  The reference's host helpers
// for the TF discretisation are compiled in place (oracle/ref_host) and compared at build time.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "dvr_oracle.h"

namespace {

struct V3
{
  float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 v3(const float *p) { return {p[0], p[1], p[2]}; }
inline float cmax(V3 a) { return std::fmax(std::fmax(a.x, a.y), a.z); }
inline float cmin(V3 a) { return std::fmin(std::fmin(a.x, a.y), a.z); }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normalize(V3 v)
{
  const float d = v.x * v.x + v.y * v.y + v.z * v.z;
  return v * (1.0f / std::sqrt(d));
}

// ---- cuRAND Philox4x32-10 ------------------------------------------------------------------------
struct Philox
{
  uint32_t key[2], ctr[4], out[4], state;

  static void round1(uint32_t c[4], const uint32_t k[2])
  {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n[4] = {hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0};
    std::memcpy(c, n, sizeof(n));
  }
  void generate()
  {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
      round1(c, k);
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    std::memcpy(out, c, sizeof(c));
  }
  void incr(uint64_t n)
  { // Philox_State_Incr(s, n)
    uint32_t nlo = (uint32_t)n, nhi = (uint32_t)(n >> 32);
    ctr[0] += nlo;
    if (ctr[0] < nlo)
      nhi++;
    ctr[1] += nhi;
    if (nhi <= ctr[1])
      return;
    if (++ctr[2])
      return;
    ++ctr[3];
  }
  void init(uint64_t seed, uint64_t subsequence, uint64_t offset)
  {
    ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    state = 0;
    // skipahead_sequence
    {
      uint32_t nlo = (uint32_t)subsequence, nhi = (uint32_t)(subsequence >> 32);
      ctr[2] += nlo;
      if (ctr[2] < nlo)
        nhi++;
      ctr[3] += nhi;
    }
    // skipahead
    state += (uint32_t)(offset & 3);
    uint64_t n = offset / 4;
    if (state > 3) {
      n += 1;
      state -= 4;
    }
    incr(n);
    generate();
  }
  uint32_t next()
  {
    const uint32_t r = out[state++];
    if (state == 4) {
      incr(1);
      generate();
      state = 0;
    }
    return r;
  }
  static float toUniform(uint32_t x) { return std::fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f); }
  float uniform() { return toUniform(next()); }
  void uniform4(float r[4])
  { // curand4: four consecutive outputs whatever the alignment
    for (int i = 0; i < 4; ++i)
      r[i] = toUniform(next());
  }
};

// ---- software texture unit (B200 measured model; profiles/texture_unit_model.md) ----------------------
// normalised coordinate u on an axis of N texels, clamp addressing, linear filter:
//   xB = u*N - 0.5 (fp32) -> 1.8 fixed point q = floor(xB*256 + 0.5), clamped to [0,(N-1)*256]
//   i = q>>8, k = q&255, taps i and min(i+1, N-1)
inline void texAxis(float u, int N, int &i0, int &i1, int &k)
{
  const float xb = u * (float)N - 0.5f;
  double qd = std::floor((double)xb * 256.0 + 0.5);
  const double qmax = (double)(N - 1) * 256.0;
  if (!(qd >= 0.0))
    qd = 0.0; // also catches NaN
  if (qd > qmax)
    qd = qmax;
  const long long q = (long long)qd;
  i0 = (int)(q >> 8);
  k = (int)(q & 255);
  i1 = std::min(i0 + 1, N - 1);
}

// ---- NanoVDB float grids: SpatialFieldSampler<nanovdb::Grid<NanoTree<float>>>, sampleSpatialField.h:80-109 ----
// Restated against the published NanoVDB 32.7 binary layout (the reference's vendored headers are not
// linked here; layout constants are static_assert-checked in oracle/ref_host/ref_host.cpp):
//   GridData 672 B (Map at 296: mInvMatF +36, mVecF +72), TreeData 64 B (mNodeOffset[3] = root), RootData<float>
//   64 B + 32-byte tiles {key u64, child i64, state u32, value f32}; upper 32^3 (child mask +4128, table +8256),
//   lower 16^3 (child mask +544, table +1088), leaf 8^3 (values +96).
struct Nvdb
{
  const uint8_t *root = nullptr;
  uint32_t tiles = 0;
  float background = 0.f;
  float invMat[9], vec[3];
  int codecLog2Bits = 0; // quantised grids: log2(bits per code); -1 = FpN (per leaf, mFlags >> 5)
  bool quant = false;

  template <typename T>
  static T rd(const uint8_t *p)
  {
    T v;
    std::memcpy(&v, p, sizeof(T));
    return v;
  }
  void open(const uint8_t *blob)
  {
    const int64_t rootOff = 672 + rd<int64_t>(blob + 672 + 24);
    root = blob + rootOff;
    tiles = rd<uint32_t>(root + 24);
    background = rd<float>(root + 28);
    for (int i = 0; i < 9; ++i)
      invMat[i] = rd<float>(blob + 296 + 36 + 4 * i);
    for (int i = 0; i < 3; ++i)
      vec[i] = rd<float>(blob + 296 + 72 + 4 * i);
    // GridData::mGridType (NanoVDB.h:220-235): Float 1, Fp4 13, Fp8 14, Fp16 15, FpN 16
    switch (rd<uint32_t>(blob + 636)) {
    case 13: quant = true; codecLog2Bits = 2; break;
    case 14: quant = true; codecLog2Bits = 3; break;
    case 15: quant = true; codecLog2Bits = 4; break;
    case 16: quant = true; codecLog2Bits = -1; break;
    default: quant = false; break;
    }
  }
  // LeafData<Fp4|Fp8|Fp16|FpN>::getValue (NanoVDB.h:3897-4040): code * mQuantum + mMinimum, evaluated as one
  // fused multiply-add like the device code nvcc generates for the reference
  float leafValue(const uint8_t *leaf, uint32_t n) const
  {
    if (!quant)
      return rd<float>(leaf + 96 + 4 * (size_t)n);
    const int b = codecLog2Bits >= 0 ? codecLog2Bits : (int)(leaf[15] >> 5);
    uint32_t code = rd<uint32_t>(leaf + 96 + 4 * (size_t)(n >> (5 - b)));
    code >>= (n & ((32u >> b) - 1u)) << b;
    code &= (1u << (1u << b)) - 1u;
    return std::fmaf((float)code, rd<float>(leaf + 84), rd<float>(leaf + 80));
  }
  static bool maskOn(const uint8_t *mask, uint32_t n) { return (rd<uint64_t>(mask + 8 * (n >> 6)) >> (n & 63)) & 1; }
  // Tree::getValue -> RootNode::getValue -> InternalNode::getValue -> LeafNode::getValue (NanoVDB.h:2947-2953,3528-3532)
  float getValue(int x, int y, int z) const
  {
    const uint64_t key = (uint64_t)((uint32_t)z >> 12) | ((uint64_t)((uint32_t)y >> 12) << 21)
        | ((uint64_t)((uint32_t)x >> 12) << 42);
    const uint8_t *tile = nullptr;
    for (uint32_t i = 0; i < tiles; ++i)
      if (rd<uint64_t>(root + 64 + 32 * (size_t)i) == key) {
        tile = root + 64 + 32 * (size_t)i;
        break;
      }
    if (!tile)
      return background;
    const int64_t child = rd<int64_t>(tile + 8);
    if (child == 0)
      return rd<float>(tile + 20);
    const uint8_t *upper = root + child;
    uint32_t n = (uint32_t)((((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7));
    if (!maskOn(upper + 32 + 4096, n))
      return rd<float>(upper + 8256 + 8 * (size_t)n);
    const uint8_t *lower = upper + rd<int64_t>(upper + 8256 + 8 * (size_t)n);
    n = (uint32_t)((((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3));
    if (!maskOn(lower + 32 + 512, n))
      return rd<float>(lower + 1088 + 8 * (size_t)n);
    const uint8_t *leaf = lower + rd<int64_t>(lower + 1088 + 8 * (size_t)n);
    return leafValue(leaf, (uint32_t)(((x & 7) << 6) | ((y & 7) << 3) | (z & 7)));
  }
  // worldToIndexF(Vec3d) = matMult(mInvMatF, xyz - mVecF): subtraction in double, fmaf chain in float (Math.h:905-910)
  float sample(V3 p) const
  {
    const float fx = (float)((double)p.x - (double)vec[0]), fy = (float)((double)p.y - (double)vec[1]),
                fz = (float)((double)p.z - (double)vec[2]);
    const float ix = std::fmaf(fx, invMat[0], std::fmaf(fy, invMat[1], fz * invMat[2]));
    const float iy = std::fmaf(fx, invMat[3], std::fmaf(fy, invMat[4], fz * invMat[5]));
    const float iz = std::fmaf(fx, invMat[6], std::fmaf(fy, invMat[7], fz * invMat[8]));
    // SampleFromVoxels<Acc,1>: ijk = floor, uvw = xyz - ijk (double), lerp a + float(w)*(b-a), z innermost
    const double dx = ix, dy = iy, dz = iz;
    const int i = (int)std::floor(dx), j = (int)std::floor(dy), k = (int)std::floor(dz);
    const float u = (float)(dx - i), v = (float)(dy - j), w = (float)(dz - k);
    auto lerp = [](float a, float b, float t) { return a + t * (b - a); };
    const float v000 = getValue(i, j, k), v001 = getValue(i, j, k + 1), v011 = getValue(i, j + 1, k + 1),
                v010 = getValue(i, j + 1, k), v100 = getValue(i + 1, j, k), v101 = getValue(i + 1, j, k + 1),
                v111 = getValue(i + 1, j + 1, k + 1), v110 = getValue(i + 1, j + 1, k);
    return lerp(lerp(lerp(v000, v001, w), lerp(v010, v011, w), v), lerp(lerp(v100, v101, w), lerp(v110, v111, w), v), u);
  }
};

struct Field
{
  const float *vox;
  int nx, ny, nz;
  Nvdb nvdb;
  bool isNvdb = false;
  V3 origin, spacing, invSpacing, lo, hi;
  float stepSize;
  bool nearest;

  float at(int x, int y, int z) const { return vox[((size_t)z * ny + y) * nx + x]; }

  // tex3D<float>, normalised coords, clamp.  Tap weights in 1/256 units (exact rule measured with
  // one-hot textures): slice weights (256-kz, kz); within a slice of weight B:
  //   X1 = floor(B*kx/256 + .5), X0 = B - X1
  //   w(x1,y1) = floor(X1*ky/256 + .5)           w(x1,y0) = X1 - w(x1,y1)
  //   w(x0,y1) = ceil (X0*ky/256 - .5)           w(x0,y0) = X0 - w(x0,y1)
  float tex(float u, float v, float w) const
  {
    if (nearest) {
      auto pick = [](float c, int N) {
        const float x = std::floor(c * (float)N);
        return (int)std::min(std::max(x, 0.0f), (float)(N - 1));
      };
      return at(pick(u, nx), pick(v, ny), pick(w, nz));
    }
    int x0, x1, kx, y0, y1, ky, z0, z1, kz;
    texAxis(u, nx, x0, x1, kx);
    texAxis(v, ny, y0, y1, ky);
    texAxis(w, nz, z0, z1, kz);
    double acc = 0.0;
    const int B[2] = {256 - kz, kz};
    const int zz[2] = {z0, z1};
    for (int s = 0; s < 2; ++s) {
      if (B[s] == 0)
        continue;
      const int X1 = (B[s] * kx + 128) >> 8;
      const int X0 = B[s] - X1;
      const int w11 = (X1 * ky + 128) >> 8;
      const int w10 = X1 - w11;
      const int w01 = (X0 * ky + 127) >> 8; // ceil(a/256 - 1/2) for integer a
      const int w00 = X0 - w01;
      if (w00) acc += (double)at(x0, y0, zz[s]) * w00;
      if (w10) acc += (double)at(x1, y0, zz[s]) * w10;
      if (w01) acc += (double)at(x0, y1, zz[s]) * w01;
      if (w11) acc += (double)at(x1, y1, zz[s]) * w11;
    }
    return (float)(acc / 256.0);
  }

  // SpatialFieldSampler<cudaTextureObject_t>::operator(), sampleSpatialField.h:66-71
  float sample(V3 p) const
  {
    if (isNvdb)
      return nvdb.sample(p);
    const V3 tc = ((p - origin) + 0.5f * spacing) * invSpacing;
    return tex(tc.x, tc.y, tc.z);
  }
};

// tex1D<float4> on the 256-texel TF (clamp, normalised, linear)
inline void tfFetch(const float *tf, float coord, float out[4])
{
  int i0, i1, k;
  texAxis(coord, DVR_TF_SIZE, i0, i1, k);
  for (int c = 0; c < 4; ++c)
    out[c] = (float)(((double)tf[4 * i0 + c] * (256 - k) + (double)tf[4 * i1 + c] * k) / 256.0);
}

// getBackgroundImage, gpu/gpu_util.h:289-296: the renderer's constant colour, or tex2D<float4> on the RGBA8
// background texture (clamp, normalised coordinates, linear, cudaReadModeNormalizedFloat; Renderer.cpp:172-179,
// utility/CudaImageTexture.cpp:139-226,316-345).  The filter is the measured B200 rule of Field::tex for one slice
// (B = 256).  Channels the texture does not have read as 0, alpha included (measured on B200: a two-channel RG8
// texture fetched as float4 returns (r, g, 0, 0); golden scene bgimage_u16x2_checkerboard_p5).
struct BgImage
{
  std::vector<uint8_t> texels; // nc bytes per texel, row-major; empty = constant colour
  int nc = 0, w = 0, h = 0;
};
static BgImage g_bgImage;

inline void backgroundAt(const DvrFrameParams &P, float sx, float sy, float out[4])
{
  const BgImage &B = g_bgImage;
  if (B.texels.empty()) {
    std::memcpy(out, P.background, 4 * sizeof(float));
    return;
  }
  int x0, x1, kx, y0, y1, ky;
  texAxis(sx, B.w, x0, x1, kx);
  texAxis(sy, B.h, y0, y1, ky);
  const int X1 = kx, X0 = 256 - kx;
  const int w11 = (X1 * ky + 128) >> 8, w10 = X1 - w11;
  const int w01 = (X0 * ky + 127) >> 8, w00 = X0 - w01;
  out[0] = out[1] = out[2] = out[3] = 0.f;
  for (int c = 0; c < B.nc; ++c) {
    auto at = [&](int x, int y) { return (double)((float)B.texels[((size_t)y * B.w + x) * B.nc + c] / 255.f); };
    out[c] = (float)((at(x0, y0) * w00 + at(x1, y0) * w10 + at(x0, y1) * w01 + at(x1, y1) * w11) / 256.0);
  }
}

inline float position(float v, float lo, float hi)
{ // gpu_math.h:176-180
  v = std::fmax(lo, std::fmin(v, hi));
  return (v - lo) * (1.f / (hi - lo));
}

struct Volume
{
  Field f;
  const float *tf;
  float vrLo, vrHi, oneOverUnitDistance;
  uint32_t id, instId;
  float xfm[12];
  bool identity;
  int zOwnBegin, zOwnEnd; // sort-last ownership test (whole volume: 0..nz)
};

struct Camera
{
  int type;
  float region[4];
  V3 pos, dir, du, dv, p00;
  float scaledAperture, aspect;
};

inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

void cameraCreateRay(const Camera &c, float sx, float sy, float rz, float rw, V3 &org, V3 &dir)
{
  sx = mixf(c.region[0], c.region[2], sx);
  sy = mixf(c.region[1], c.region[3], sy);
  if (c.type == DVR_CAMERA_PERSPECTIVE) {
    org = c.pos;
    dir = c.p00 + sx * c.du + sy * c.dv;
    if (c.scaledAperture > 0.f) {
      const float r = std::sqrt(rz) * c.scaledAperture;
      const float phi = 2.f * float(M_PI) * rw;
      const float lx = r * std::cos(phi), ly = r * std::sin(phi);
      const V3 lp = (lx * c.du) + ((ly * c.aspect) * c.dv);
      org = org + lp;
      dir = dir - lp;
    }
    dir = normalize(dir);
  } else {
    dir = c.dir;
    org = c.p00 + sx * c.du + sy * c.dv;
  }
}

// Intersectors_ptx.cu:248-274 (plus the AABB/interval overlap the BVH traversal implies)
bool intersectVolume(const Volume &v, V3 org, V3 dir, float tmin, float tmax, float &t0, float &t1)
{
  const V3 inv = {1.f / dir.x, 1.f / dir.y, 1.f / dir.z};
  const V3 mins = (v.f.lo - org) * inv;
  const V3 maxs = (v.f.hi - org) * inv;
  const V3 nears = {std::fmin(mins.x, maxs.x), std::fmin(mins.y, maxs.y), std::fmin(mins.z, maxs.z)};
  const V3 fars = {std::fmax(mins.x, maxs.x), std::fmax(mins.y, maxs.y), std::fmax(mins.z, maxs.z)};
  const float tn = cmax(nears), tf = cmin(fars);
  if (!(tn < tf))
    return false;
  if (tf < tmin || tn > tmax)
    return false;
  t0 = std::fmax(tmin, std::fmin(tn, tmax));
  t1 = std::fmax(tmin, std::fmin(tf, tmax));
  return true;
}

// _rayMarchVolume, volumeIntegration.h:64-103
void marchSegment(const Volume &v, V3 org, V3 dir, float lower, float upper, float invSamplingRate, Philox &rng,
    V3 &color, float &opacity, uint64_t &samples)
{
  const float stepSize = v.f.stepSize * invSamplingRate;
  const float exponent = stepSize * v.oneOverUnitDistance;
  lower += stepSize * rng.uniform();
  float transmittance = 1.f;
  while (opacity < 0.99f && (upper - lower) >= 0.f) {
    const V3 p = org + dir * lower;
    bool own = true;
    if (v.zOwnBegin > 0 || v.zOwnEnd < v.f.nz) {
      const V3 tc = ((p - v.f.origin) + 0.5f * v.f.spacing) * v.f.invSpacing;
      int zc = (int)std::floor(tc.z * (float)v.f.nz - 0.5f);
      zc = std::min(std::max(zc, 0), v.f.nz - 1);
      own = zc >= v.zOwnBegin && zc < v.zOwnEnd;
    }
    if (own) {
      const float s = v.f.sample(p);
      samples++;
      if (!std::isnan(s)) {
        float co[4];
        tfFetch(v.tf, position(s, v.vrLo, v.vrHi), co);
        const float stepTransmittance = std::pow(1.f - co[3], exponent);
        const float w = transmittance * (1.f - stepTransmittance);
        color.x += w * co[0];
        color.y += w * co[1];
        color.z += w * co[2];
        opacity += w;
        transmittance *= stepTransmittance;
      }
    }
    lower += stepSize;
  }
}

inline V3 xfmPoint(const float *m, V3 p)
{
  return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
      m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
inline V3 xfmVector(const float *m, V3 v)
{
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
      m[8] * v.x + m[9] * v.y + m[10] * v.z};
}

// rayMarchAllVolumes, volumeIntegration.h:317-350
float rayMarchAllVolumes(const std::vector<Volume> &vols, V3 org, V3 dir, float tfar, float invSamplingRate,
    Philox &rng, V3 &color, float &opacity, uint32_t &objID, uint32_t &instID, uint64_t &samples)
{
  float rayLower = 0.f;
  const float rayUpper = tfar;
  float depth = tfar;
  bool firstHit = true;
  int last = -1;
  do {
    int best = -1;
    float bt0 = 0.f, bt1 = 0.f;
    V3 bo = org, bd = dir;
    for (int i = 0; i < (int)vols.size(); ++i) {
      if (i == last)
        continue;
      V3 lo = org, ld = dir;
      if (!vols[i].identity) {
        lo = xfmPoint(vols[i].xfm, org);
        ld = xfmVector(vols[i].xfm, dir);
      }
      float t0, t1;
      if (!intersectVolume(vols[i], lo, ld, rayLower, rayUpper, t0, t1))
        continue;
      if (best < 0 || t0 < bt0) {
        best = i;
        bt0 = t0;
        bt1 = t1;
        bo = lo;
        bd = ld;
      }
    }
    if (best < 0)
      break;
    const Volume &v = vols[best];
    if (firstHit) {
      objID = v.id;
      instID = v.instId;
      firstHit = false;
    }
    depth = std::fmin(depth, bt0);
    bt1 = std::fmin(tfar, bt1);
    const float start = bt0 + v.f.stepSize * rng.uniform(); // volumeIntegration.h:117-120
    marchSegment(v, bo, bd, start, bt1, invSamplingRate, rng, color, opacity, samples);
    rayLower = bt1 + 1e-3f;
    last = best;
  } while (opacity < 0.99f);
  return depth;
}

// ---- dpt renderer: delta (Woodcock) tracking -----------------------------------------------------------
// The grid: the reference's geometry (UniformGrid::init, UniformGrid.cu:152-154: ceil(dims/16) cells dividing
// the field bounds evenly).  Its content is NOT the reference's build (SURVEY quirks Q7/Q8 make that one
// non-conservative, i.e. the reference's own images are biased); it is the conservative build the product
// documents in DESIGN.md: per cell the min/max over every voxel a trilinear stencil inside the cell can touch
// (one voxel of margin), majorant = max TF alpha over the texels those values can address.
struct DdaGrid
{
  int gx = 0, gy = 0, gz = 0;
  std::vector<float> maj;
};

DdaGrid buildDdaGrid(const Volume &v)
{
  DdaGrid g;
  int nx = v.f.nx, ny = v.f.ny, nz = v.f.nz, bx = 0, by = 0, bz = 0;
  float sx, sy, sz; // voxel units spanned by the bounds
  if (v.f.isNvdb) {
    int32_t bb[6];
    std::memcpy(bb, v.f.nvdb.root, sizeof(bb)); // RootData::mBBox (index space)
    bx = bb[0]; by = bb[1]; bz = bb[2];
    nx = std::max(bb[3] - bb[0] + 1, 1); ny = std::max(bb[4] - bb[1] + 1, 1); nz = std::max(bb[5] - bb[2] + 1, 1);
    sx = (float)nx; sy = (float)ny; sz = (float)nz;
  } else {
    sx = (float)nx - 1.f; sy = (float)ny - 1.f; sz = (float)nz - 1.f;
  }
  g.gx = (nx + 15) / 16; g.gy = (ny + 15) / 16; g.gz = (nz + 15) / 16;
  const float wx = sx / (float)g.gx, wy = sy / (float)g.gy, wz = sz / (float)g.gz;
  g.maj.assign((size_t)g.gx * g.gy * g.gz, 0.f);
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int cz = 0; cz < g.gz; ++cz)
    for (int cy = 0; cy < g.gy; ++cy)
      for (int cx = 0; cx < g.gx; ++cx) {
        int x0 = (int)std::floor(cx * wx) - 1, x1 = (int)std::ceil((cx + 1) * wx) + 1;
        int y0 = (int)std::floor(cy * wy) - 1, y1 = (int)std::ceil((cy + 1) * wy) + 1;
        int z0 = (int)std::floor(cz * wz) - 1, z1 = (int)std::ceil((cz + 1) * wz) + 1;
        if (!v.f.isNvdb) {
          x0 = std::max(x0, 0); y0 = std::max(y0, 0); z0 = std::max(z0, 0);
          x1 = std::min(x1, nx - 1); y1 = std::min(y1, ny - 1); z1 = std::min(z1, nz - 1);
        }
        float lo = std::numeric_limits<float>::max(), hi = -std::numeric_limits<float>::max();
        for (int z = z0; z <= z1; ++z)
          for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
              const float s = v.f.isNvdb ? v.f.nvdb.getValue(x + bx, y + by, z + bz) : v.f.at(x, y, z);
              lo = std::fmin(lo, s);
              hi = std::fmax(hi, s);
            }
        float m = 0.f;
        if (lo <= hi) {
          const float c0 = position(lo, v.vrLo, v.vrHi), c1 = position(hi, v.vrLo, v.vrHi);
          int i0 = (int)std::floor(c0 * 256.0f - 0.5f) - 1, i1 = (int)std::floor(c1 * 256.0f - 0.5f) + 2;
          i0 = std::max(0, std::min(i0, DVR_TF_SIZE - 1));
          i1 = std::max(0, std::min(i1, DVR_TF_SIZE - 1));
          for (int i = i0; i <= i1; ++i)
            m = std::fmax(m, v.tf[4 * i + 3]);
        }
        g.maj[((size_t)cz * g.gy + cy) * g.gx + cx] = m;
      }
  return g;
}

// _sampleDistance (volumeIntegration.h:167-238) + dda3 (dda.h:43-121) + projectOnGrid (uniformGrid.h:44-50)
float sampleDistanceSegment(const Volume &v, const DdaGrid &g, V3 lorg, V3 ldir, float tLower, float tUpper,
    Philox &rng, V3 &albedo, float &extinction, float &tr, uint64_t &samples)
{
  const float stepSize = v.f.stepSize;
  float t_out = tUpper;
  tr = 1.f;
  const V3 oorg = lorg + ldir * tLower;
  const float rayUpper = tUpper - tLower;
  const V3 rcp = {ldir.x != 0.f ? 1.f / ldir.x : 0.f, ldir.y != 0.f ? 1.f / ldir.y : 0.f,
      ldir.z != 0.f ? 1.f / ldir.z : 0.f};
  const V3 lo = (v.f.lo - oorg) * rcp, hi = (v.f.hi - oorg) * rcp;
  const V3 tnear = {std::fmin(lo.x, hi.x), std::fmin(lo.y, hi.y), std::fmin(lo.z, hi.z)};
  const V3 tfar = {std::fmax(lo.x, hi.x), std::fmax(lo.y, hi.y), std::fmax(lo.z, hi.z)};
  const V3 v01 = {(oorg.x - v.f.lo.x) / (v.f.hi.x - v.f.lo.x), (oorg.y - v.f.lo.y) / (v.f.hi.y - v.f.lo.y),
      (oorg.z - v.f.lo.z) / (v.f.hi.z - v.f.lo.z)};
  int c[3] = {std::min(std::max((int)(v01.x * (float)g.gx), 0), g.gx - 1),
      std::min(std::max((int)(v01.y * (float)g.gy), 0), g.gy - 1),
      std::min(std::max((int)(v01.z * (float)g.gz), 0), g.gz - 1)};
  const int gd[3] = {g.gx, g.gy, g.gz};
  const float d[3] = {ldir.x, ldir.y, ldir.z};
  const float tn[3] = {tnear.x, tnear.y, tnear.z}, tf_[3] = {tfar.x, tfar.y, tfar.z};
  float dist[3], tnext[3];
  int step[3], stop[3];
  for (int a = 0; a < 3; ++a) {
    dist[a] = (tf_[a] - tn[a]) / (float)gd[a];
    step[a] = d[a] > 0.f ? 1 : -1;
    stop[a] = d[a] > 0.f ? gd[a] : -1;
    tnext[a] = std::fmaf((float)(d[a] > 0.f ? c[a] + 1 : gd[a] - c[a]), dist[a], tn[a]);
  }
  float t0 = 0.f;
  while (true) {
    const float tmin3 = std::fmin(std::fmin(tnext[0], tnext[1]), tnext[2]);
    const float t1 = std::fmin(tmin3, rayUpper);
    const float majorant = g.maj[((size_t)c[2] * g.gy + c[1]) * g.gx + c[0]];
    float t = t0;
    while (majorant > 0.f) {
      t = std::fmaf(-(std::log(1.f - rng.uniform()) / majorant), stepSize, t);
      if (t >= t1)
        break;
      const V3 p = lorg + ldir * (t + tLower);
      const float s = v.f.sample(p);
      samples++;
      if (!std::isnan(s)) {
        float co[4];
        tfFetch(v.tf, position(s, v.vrLo, v.vrHi), co);
        albedo = {co[0], co[1], co[2]};
        extinction = co[3];
        const float u = rng.uniform();
        if (extinction >= u * majorant) {
          tr = 0.f;
          t_out = t;
          return t_out + tLower;
        }
      }
    }
    bool out = false;
    for (int a = 0; a < 3 && !out; ++a)
      if (tnext[a] == tmin3) {
        tnext[a] += dist[a];
        c[a] += step[a];
        if (c[a] == stop[a])
          out = true;
      }
    if (out)
      break;
    t0 = t1;
  }
  return t_out + tLower;
}

// sampleDistanceAllVolumes, volumeIntegration.h:352-389
float sampleDistanceAllVolumes(const std::vector<Volume> &vols, const std::vector<DdaGrid> &grids, V3 org, V3 dir,
    float tmin, float tfar, Philox &rng, V3 &albedo, float &extinction, float &transmittance, uint64_t &samples)
{
  float rayLower = tmin;
  float depth = tfar;
  transmittance = 1.f;
  int last = -1;
  while (true) {
    int best = -1;
    float bt0 = 0.f, bt1 = 0.f;
    V3 bo = org, bd = dir;
    for (int i = 0; i < (int)vols.size(); ++i) {
      if (i == last)
        continue;
      V3 lo = org, ld = dir;
      if (!vols[i].identity) {
        lo = xfmPoint(vols[i].xfm, org);
        ld = xfmVector(vols[i].xfm, dir);
      }
      float t0, t1;
      if (!intersectVolume(vols[i], lo, ld, rayLower, tfar, t0, t1))
        continue;
      if (best < 0 || t0 < bt0) {
        best = i;
        bt0 = t0;
        bt1 = t1;
        bo = lo;
        bd = ld;
      }
    }
    if (best < 0)
      break;
    bt1 = std::fmin(tfar, bt1);
    V3 alb{0.f, 0.f, 0.f};
    float ext = 0.f, tr = 0.f;
    const float dd = sampleDistanceSegment(vols[best], grids[best], bo, bd, bt0, bt1, rng, alb, ext, tr, samples);
    if (dd < depth) {
      depth = dd;
      albedo = alb;
      extinction = ext;
      transmittance = tr;
    }
    rayLower = bt1 + 1e-3f;
    last = best;
  }
  return depth;
}

// computeOrthonormalBasis / sampleUnitSphere, gpu_util.h:205-243
V3 sampleUnitSphere(Philox &rng, V3 n)
{
  const float cost = 1.f - 2.f * rng.uniform();
  const float sint = std::sqrt(std::fmax(0.f, 1.f - cost * cost));
  const float phi = 2.f * 3.14159265358979323846f * rng.uniform();
  const float sign = n.z >= 0.0f ? 1.0f : -1.0f;
  const float a = -1.0f / (sign + n.z);
  const float b = n.x * n.y * a;
  const V3 u = {1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x};
  const V3 w = {b, sign + n.y * n.y * a, -n.y};
  const float sx = sint * std::cos(phi), sy = sint * std::sin(phi), sz = -cost;
  return {u.x * sx + w.x * sy + n.x * sz, u.y * sx + w.y * sy + n.y * sz, u.z * sx + w.z * sy + n.z * sz};
}

struct DptPath // PathData, DiffusePathTracer_ptx.cu:45-50 (declared OUTSIDE the numIterations loop)
{
  int depth = 0;
  V3 Lw{1.f, 1.f, 1.f};
};

// the bounce loop of DiffusePathTracer_ptx.cu:111-196 for a world without surfaces
V3 dptTracePath(const float bg[4], const std::vector<Volume> &vols, const std::vector<DdaGrid> &grids, V3 org, V3 dir,
    const DvrFrameParams &P, Philox &rng, DptPath &path, uint64_t &samples)
{
  const int maxDepth = P.maxDepth <= 0 ? 5 : std::min(P.maxDepth, 256);
  const float occl = P.occlusionDistance > 0.f ? P.occlusionDistance : 1e20f;
  float tmin = 0.f, tmax = std::numeric_limits<float>::max();
  while (true) {
    V3 volumeColor{0.f, 0.f, 0.f};
    float volumeOpacity = 0.f, Tr = 0.f;
    const float volumeDepth =
        sampleDistanceAllVolumes(vols, grids, org, dir, tmin, tmax, rng, volumeColor, volumeOpacity, Tr, samples);
    if (!(Tr < 1.f))
      break;
    if (path.depth++ >= maxDepth) {
      path.Lw = {0.f, 0.f, 0.f};
      break;
    }
    const V3 pos = org + dir * volumeDepth;
    path.Lw = path.Lw * volumeColor;
    const float Pr = cmax(path.Lw);
    if (Pr < .2f) {
      if (rng.uniform() > Pr) {
        path.Lw = {0.f, 0.f, 0.f};
        break;
      }
      path.Lw = {path.Lw.x / Pr, path.Lw.y / Pr, path.Lw.z / Pr};
    }
    const V3 scatter = sampleUnitSphere(rng, V3{-dir.x, -dir.y, -dir.z});
    org = pos;
    dir = scatter;
    tmin = 0.f;
    tmax = occl;
  }
  if (path.depth)
    return path.Lw * P.ambientRadiance;
  return {bg[0], bg[1], bg[2]};
}

inline float clamp01(float v) { return std::fmin(std::fmax(v, 0.f), 1.f); }
inline float toSrgb(float c)
{
  c = clamp01(c);
  const float hi = std::pow(c, 0.41666f) * 1.055f - 0.055f;
  const float lo = c * 12.92f;
  return c < 0.0031308f ? lo : hi;
}
inline uint32_t packUnorm4x8(float r, float g, float b, float a)
{
  const uint32_t R = (uint8_t)std::round(clamp01(r) * 255.f), G = (uint8_t)std::round(clamp01(g) * 255.f),
                 B = (uint8_t)std::round(clamp01(b) * 255.f), A = (uint8_t)std::round(clamp01(a) * 255.f);
  return R | (G << 8) | (B << 16) | (A << 24);
}

struct Frame
{
  const DvrFrameParams *p;
  float *accum;
  void *color;
  float *depth;
  uint32_t *prim, *obj, *inst;
  float *albedo, *normal;
};

void writeOutputColor(const Frame &f, const float acc[4], uint32_t idx, int frameIDplusOffset)
{ // gpu_util.h:375-391
  const float div = float(frameIDplusOffset + 1);
  float c[4] = {acc[0] / div, acc[1] / div, acc[2] / div, acc[3] / div};
  const float m = std::fmax(1e-12f, 1.f - std::fmax(std::fmax(c[0], c[1]), c[2]));
  c[0] /= m;
  c[1] /= m;
  c[2] /= m;
  if (f.p->format == DVR_FORMAT_UFIXED8_RGBA_SRGB)
    ((uint32_t *)f.color)[idx] = packUnorm4x8(toSrgb(c[0]), toSrgb(c[1]), toSrgb(c[2]), c[3]);
  else if (f.p->format == DVR_FORMAT_UFIXED8_VEC4)
    ((uint32_t *)f.color)[idx] = packUnorm4x8(c[0], c[1], c[2], c[3]);
  else
    std::memcpy((float *)f.color + 4 * (size_t)idx, c, sizeof(c));
}

void accumResults(const Frame &f, uint32_t px, uint32_t py, const float color[4], float depth, const float albedo[3],
    const float normal[3], uint32_t primID, uint32_t objID, uint32_t instID, int frameIDOffset)
{ // gpu_util.h:393-443
  const DvrFrameParams &P = *f.p;
  const uint32_t idx = px + py * P.width;
  const int frameID = P.frameID + frameIDOffset;
  const float m = 1.f + std::fmax(0.f, std::fmax(std::fmax(color[0], color[1]), color[2]));
  float *acc = f.accum + 4 * (size_t)idx;
  acc[0] += color[0] / m;
  acc[1] += color[1] / m;
  acc[2] += color[2] / m;
  acc[3] += color[3];
  if (f.albedo)
    for (int c = 0; c < 3; ++c)
      f.albedo[3 * (size_t)idx + c] += albedo[c];
  if (f.normal)
    for (int c = 0; c < 3; ++c)
      f.normal[3 * (size_t)idx + c] += normal[c];
  bool closer = true;
  if (f.depth) {
    closer = depth < f.depth[idx];
    if (closer)
      f.depth[idx] = depth;
  }
  if (closer) {
    if (f.prim) f.prim[idx] = primID;
    if (f.obj) f.obj[idx] = objID;
    if (f.inst) f.inst[idx] = instID;
  }
  writeOutputColor(f, acc, idx, frameID);
  if (P.checkerboardID == 0 && frameID == 0) {
    const uint32_t adj[3][2] = {{px + 1, py}, {px, py + 1}, {px + 1, py + 1}};
    for (auto &a : adj)
      if (a[0] < P.width && a[1] < P.height)
        writeOutputColor(f, acc, a[0] + a[1] * P.width, frameID);
  }
}

void makeVolume(const OracleVolume &o, Volume &v)
{
  v.f.vox = o.voxels;
  v.f.nx = o.dims[0];
  v.f.ny = o.dims[1];
  v.f.nz = o.dims[2];
  v.f.origin = v3(o.origin);
  v.f.spacing = v3(o.spacing);
  v.f.invSpacing = {1.f / (o.spacing[0] * (float)o.dims[0]), 1.f / (o.spacing[1] * (float)o.dims[1]),
      1.f / (o.spacing[2] * (float)o.dims[2])};
  v.f.lo = v.f.origin;
  v.f.hi = {o.origin[0] + ((float)o.dims[0] - 1.f) * o.spacing[0], o.origin[1] + ((float)o.dims[1] - 1.f) * o.spacing[1],
      o.origin[2] + ((float)o.dims[2] - 1.f) * o.spacing[2]};
  v.f.stepSize = std::fmin(std::fmin(o.spacing[0] / 2.f, o.spacing[1] / 2.f), o.spacing[2] / 2.f);
  v.f.nearest = o.filterNearest != 0;
  if (o.nvdbGrid) { // NvdbRegularField: bounds = world bbox, step = min(voxelSize)/2 (NvdbRegularField.cpp:105-127)
    const uint8_t *blob = (const uint8_t *)o.nvdbGrid;
    v.f.isNvdb = true;
    v.f.nvdb.open(blob);
    double wb[6], vs[3];
    std::memcpy(wb, blob + 560, sizeof(wb));
    std::memcpy(vs, blob + 608, sizeof(vs));
    v.f.lo = {(float)wb[0], (float)wb[1], (float)wb[2]};
    v.f.hi = {(float)wb[3], (float)wb[4], (float)wb[5]};
    v.f.stepSize = std::fmin(std::fmin((float)vs[0], (float)vs[1]), (float)vs[2]) / 2.0f;
  }
  v.tf = o.tf;
  v.vrLo = o.valueRange[0];
  v.vrHi = o.valueRange[1];
  v.oneOverUnitDistance = 1.0f / o.unitDistance;
  v.id = o.id;
  v.instId = o.instanceId;
  std::memcpy(v.xfm, o.worldToObject, sizeof(v.xfm));
  static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  v.identity = std::memcmp(v.xfm, ident, sizeof(ident)) == 0;
  v.zOwnBegin = o.zOwnEnd > o.zOwnBegin ? o.zOwnBegin : 0;
  v.zOwnEnd = o.zOwnEnd > o.zOwnBegin ? o.zOwnEnd : o.dims[2];
}

} // namespace

extern "C" {

float oracle_nvdb_sample(const void *grid, float x, float y, float z)
{
  Nvdb n;
  n.open((const uint8_t *)grid);
  return n.sample({x, y, z});
}

float oracle_tex3d(const float *voxels, const int dims[3], float u, float v, float w)
{
  Field f{};
  f.vox = voxels;
  f.nx = dims[0];
  f.ny = dims[1];
  f.nz = dims[2];
  f.nearest = false;
  return f.tex(u, v, w);
}

void oracle_tex1d_tf(const float *tf, float coord, float out[4]) { tfFetch(tf, coord, out); }

// raw Philox4x32-10 block function (Random123 known-answer vectors)
void oracle_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  Philox r;
  std::memcpy(r.ctr, ctr, sizeof(r.ctr));
  std::memcpy(r.key, key, sizeof(r.key));
  r.generate();
  std::memcpy(out, r.out, sizeof(r.out));
}

void oracle_philox_uniforms(uint64_t seed, uint64_t offset, int n, float *out)
{
  Philox r;
  r.init(seed, 0, offset);
  for (int i = 0; i < n; ++i)
    out[i] = r.uniform();
}

int oracle_render(const DvrFrameParams *params, const DvrCamera *camera, const OracleVolume *volumes, int nVolumes,
    const OracleBuffers *buffers, uint64_t *samplesOut, int rowBegin, int rowEnd)
{
  if (!params || !camera || !buffers || !buffers->colorAccumulation || !buffers->outColor)
    return -1;
  const DvrFrameParams &P = *params;
  std::vector<Volume> vols(nVolumes);
  for (int i = 0; i < nVolumes; ++i)
    makeVolume(volumes[i], vols[i]);
  Camera cam;
  cam.type = camera->type;
  std::memcpy(cam.region, camera->region, sizeof(cam.region));
  cam.pos = v3(camera->pos);
  cam.dir = v3(camera->dir);
  cam.du = v3(camera->du);
  cam.dv = v3(camera->dv);
  cam.p00 = v3(camera->p00);
  cam.scaledAperture = camera->scaledAperture;
  cam.aspect = camera->aspect;

  Frame F{params, buffers->colorAccumulation, buffers->outColor, buffers->depth, buffers->primId, buffers->objId,
      buffers->instId, buffers->albedo, buffers->normal};
  const size_t npx = (size_t)P.width * P.height;

  // Frame::newFrame reset, frame/Frame.cu:590-647
  if (P.frameID == 0 && P.checkerboardID <= 0 && rowBegin <= 0) {
    std::fill(F.accum, F.accum + 4 * npx, 0.f);
    if (F.depth) std::fill(F.depth, F.depth + npx, std::numeric_limits<float>::max());
    if (F.prim) std::fill(F.prim, F.prim + npx, 0u);
    if (F.obj) std::fill(F.obj, F.obj + npx, 0u);
    if (F.inst) std::fill(F.inst, F.inst + npx, 0u);
    if (F.albedo) std::fill(F.albedo, F.albedo + 3 * npx, 0.f);
    if (F.normal) std::fill(F.normal, F.normal + 3 * npx, 0.f);
  }

  const bool cb = P.checkerboardID >= 0;
  const int launchW = cb ? (P.width + 1) / 2 : P.width, launchH = cb ? (P.height + 1) / 2 : P.height;
  const int iters = cb ? 1 : std::max(P.numIterations, 1);
  const bool centered = P.integrator == DVR_INTEGRATOR_RAYCAST;
  const float invW = 1.f / (float)P.width, invH = 1.f / (float)P.height;
  const int y0 = rowBegin > 0 ? rowBegin : 0, y1 = (rowEnd > 0 && rowEnd < launchH) ? rowEnd : launchH;
  uint64_t samples = 0;
  const bool dpt = P.integrator == DVR_INTEGRATOR_DPT;
  std::vector<DdaGrid> grids;
  if (dpt)
    for (const Volume &v : vols)
      grids.push_back(buildDdaGrid(v));

#pragma omp parallel for schedule(dynamic, 1) reduction(+ : samples)
  for (int ly = y0; ly < y1; ++ly) {
    for (int lx = 0; lx < launchW; ++lx) {
      const int x = cb ? lx * 2 + (P.checkerboardID & 1) : lx;
      const int y = cb ? ly * 2 + ((P.checkerboardID >> 1) & 1) : ly;
      if ((uint32_t)x >= P.width || (uint32_t)y >= P.height)
        continue;
      Philox rng;
      rng.init((uint64_t)(int64_t)(int)(y * (int)P.width + x), 0, (uint64_t)((int64_t)P.frameID * 512));
      DptPath path;
      for (int it = 0; it < iters; ++it) {
        float r[4];
        rng.uniform4(r);
        const float sx = (centered ? (float)x : (float)x + r[0]) * invW;
        const float sy = (centered ? (float)y : (float)y + r[1]) * invH;
        V3 org, dir;
        cameraCreateRay(cam, sx, sy, r[2], r[3], org, dir);
        if (P.integrator == DVR_INTEGRATOR_TEST) { // Test_ptx.cu:52-69
          const float c4[4] = {dir.x, dir.y, dir.z, 1.f};
          const float alb[3] = {dir.x, dir.y, dir.z};
          const float nrm[3] = {-dir.x, -dir.y, -dir.z};
          accumResults(F, (uint32_t)x, (uint32_t)y, c4, 1.f, alb, nrm, ~0u, ~0u, ~0u, 0);
          break;
        }
        if (dpt) { // DiffusePathTracer_ptx.cu:96-215: depth / ids keep their initial values (tmax, ~0u)
          const float nrm[3] = {dir.x, dir.y, dir.z};
          float bgd[4];
          backgroundAt(P, sx, sy, bgd);
          const V3 c = dptTracePath(bgd, vols, grids, org, dir, P, rng, path, samples);
          const float c4[4] = {c.x, c.y, c.z, 1.f};
          const float alb[3] = {bgd[0], bgd[1], bgd[2]};
          accumResults(F, (uint32_t)x, (uint32_t)y, c4, std::numeric_limits<float>::max(), alb, nrm, ~0u, ~0u, ~0u, it);
          continue;
        }
        V3 color{0.f, 0.f, 0.f};
        float opacity = 0.f;
        uint32_t objID = ~0u, instID = ~0u;
        const float vdepth = rayMarchAllVolumes(vols, org, dir, std::numeric_limits<float>::max(),
            P.inverseVolumeSamplingRate, rng, color, opacity, objID, instID, samples);
        const float depth = std::fmin(1e30f, vdepth);
        color = color * opacity; // Raycast_ptx.cu:159
        const float om = 1.f - opacity;
        float bg[4];
        backgroundAt(P, sx, sy, bg);
        color.x += bg[0] * om;
        color.y += bg[1] * om;
        color.z += bg[2] * om;
        opacity += bg[3] * om;
        const float c4[4] = {color.x, color.y, color.z, opacity};
        const float alb[3] = {color.x, color.y, color.z};
        const float nrm[3] = {dir.x, dir.y, dir.z};
        accumResults(F, (uint32_t)x, (uint32_t)y, c4, depth, alb, nrm, 0u, objID, instID, it);
      }
    }
  }
  if (samplesOut)
    *samplesOut = samples;
  return 0;
}

// Background image of the following oracle_render calls (test-infrastructure state, not thread safe): `texels` holds
// w*h texels of `channels` (1, 2 or 4) bytes as the renderer's RGBA8 staging pass leaves them; NULL clears it.
int oracle_set_background_image(const uint8_t *texels, int channels, int w, int h)
{
  g_bgImage = BgImage();
  if (!texels)
    return 0;
  if ((channels != 1 && channels != 2 && channels != 4) || w <= 0 || h <= 0)
    return -1;
  g_bgImage.texels.assign(texels, texels + (size_t)w * h * channels);
  g_bgImage.nc = channels;
  g_bgImage.w = w;
  g_bgImage.h = h;
  return 0;
}

// the delta-tracking grid the dpt restatement walks (dims + majorants), for checking the product's grid
int oracle_dda_majorants(const OracleVolume *volume, int32_t dims[3], float *out, size_t capacity)
{
  if (!volume || !dims)
    return -1;
  Volume v;
  makeVolume(*volume, v);
  const DdaGrid g = buildDdaGrid(v);
  dims[0] = g.gx;
  dims[1] = g.gy;
  dims[2] = g.gz;
  if (out) {
    if (capacity < g.maj.size())
      return -2;
    std::memcpy(out, g.maj.data(), g.maj.size() * sizeof(float));
  }
  return 0;
}

} // extern "C"
