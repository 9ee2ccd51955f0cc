// dvr_device.cuh — device-side building blocks of the B200 DVR path (sm_100a).
//
// Everything here is written from the arithmetic description in SURVEY.md Appendix A;
// each block cites the reference (NVIDIA/VisRTX v0.13.0) lines whose results it must
// reproduce.  No reference code is included or linked.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#include "dvr_nanovdb.cuh"

namespace dvr {

enum FieldKind : int
{
  FIELD_STRUCTURED = 0, // 3-D array texture (structuredRegular)
  FIELD_NANOVDB = 1,    // NanoVDB float grid in linear device memory ("nanovdb")
  FIELD_NANOVDB_QUANT = 2 // NanoVDB Fp4 / Fp8 / Fp16 / FpN grid (codes decoded per tap)
};

// ---------------------------------------------------------------------------------------
// POD descriptions of the scene as the kernels see it
// ---------------------------------------------------------------------------------------

struct FieldDev
{
  cudaTextureObject_t tex; // clamp / normalised / linear|point 3-D array texture
  float3 origin;
  float3 spacing;
  float3 invSpacing; // 1 / (spacing * dims)   (StructuredRegularField.cpp:188-189)
  float3 boundsLo;   // origin
  float3 boundsHi;   // origin + (dims-1)*spacing (StructuredRegularField.cpp:166-173)
  float stepSize;    // 0.5 * min(spacing)        (StructuredRegularField.cpp:175-178)
  int3 dims;         // GLOBAL voxel dims
  // sort-last slab residency (zTexBegin == 0 && texDepth == dims.z for a whole volume)
  int zOwnBegin, zOwnEnd; // slices whose samples this field owns: cell z in [zOwnBegin, zOwnEnd)
  int zTexBegin;          // global index of texture slice 0
  int texDepth;           // resident slices
  // macrocells (16^3 voxels), x-fastest
  int3 gridDims;
  const float2 *valueRanges; // (min,max) of every value a fetch inside the cell can return
  // NanoVDB fields: dims = extent of the index bounding box, voxel coordinate = index - nv.bboxMin
  int kind;
  NvdbDev nv;
};

struct VolumeDev
{
  FieldDev f;
  const float4 *tf; // DVR_TF_SIZE texels (device)
  float vrLower;    // valueRange
  float vrUpper;
  float oneOverUnitDistance;
  uint32_t id;
  const float *maxOpacities; // per macrocell: max TF alpha over the cell's range
  const float *maxOpacitiesCoarse; // per 4x4x4 block of macrocells (64^3 voxels): max of the children
  int3 coarseDims;
  const float *maxOpacitiesCoarse2; // per 4x4x4 block of those (256^3 voxels)
  int3 coarse2Dims;
  // delta-tracking grid in the reference's geometry (ceil(dims/16) cells dividing the bounds evenly)
  const float *ddaMaxOpacities;
  int3 ddaDims;
};

struct InstanceDev
{
  VolumeDev v;
  float xfm[12]; // world -> object, row-major 3x4
  uint32_t instId;
  uint32_t identity; // xfm is the identity: skip the transform (bit-identical result)
};

struct CameraDev
{
  int type; // 0 perspective, 1 orthographic
  float4 region;
  float3 pos, dir;
  float3 du, dv, p00;
  float scaledAperture, aspect;
};

struct BuffersDev
{
  float4 *accum;
  uint32_t *outU32;
  float4 *outF32;
  void *outMirror; // optional second destination of the encoded colour (pinned host memory), same layout
  float *depth;
  float *depthMirror; // optional write-only second destination of the depth channel (peer pointer, sort-last)
  uint32_t *primId, *objId, *instId;
  float *albedo, *normal; // packed vec3
};

// ---------------------------------------------------------------------------------------
// small vector helpers (no glm in the product)
// ---------------------------------------------------------------------------------------
// All arithmetic that decides WHICH samples a ray takes (ray set-up, slab test, sample positions) is
// written with explicit round-to-nearest intrinsics: __fmul_rn/__fadd_rn are never contracted and
// __fmaf_rn is a fused multiply-add, so every template instantiation of the kernels (skipping on/off,
// stats, slab) executes bit-identical arithmetic.  The FMA placement follows what nvcc's default
// contraction produces for the reference's expressions (a*b + c => fma).
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b)
{
  return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z));
}
__device__ __forceinline__ float3 operator-(float3 a, float3 b)
{
  return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z));
}
__device__ __forceinline__ float3 operator*(float3 a, float3 b)
{
  return f3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float3 operator*(float3 a, float s)
{
  return f3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s));
}
__device__ __forceinline__ float3 operator*(float s, float3 a) { return a * s; }
// a*s + b, fused
__device__ __forceinline__ float3 madd3(float3 a, float s, float3 b)
{
  return f3(__fmaf_rn(a.x, s, b.x), __fmaf_rn(a.y, s, b.y), __fmaf_rn(a.z, s, b.z));
}
__device__ __forceinline__ float max3(float3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
__device__ __forceinline__ float min3(float3 a) { return fminf(fminf(a.x, a.y), a.z); }
__device__ __forceinline__ float3 normalize3(float3 v)
{
  // glm::normalize = v * inversesqrt(dot(v,v)); glm's inversesqrt is 1/sqrt(x)
  const float d = __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmul_rn(v.x, v.x)));
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(d));
  return v * inv;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 with cuRAND's stream layout
//   curand_init(seed, 0, offset): key=(seed,0), counter += offset/4, position offset%4
//   (createScreenSample.h:58; /usr/local/cuda/include/curand_kernel.h:971-1037)
// The generator is a pure function of (key, block counter), so the state is 3 registers
// plus the 4 cached outputs instead of cuRAND's 64-byte struct.
// ---------------------------------------------------------------------------------------
struct Philox
{
  uint32_t key0, key1;
  uint32_t ctr0, ctr1; // 64-bit block counter (ctr.x, ctr.y); ctr.z/w stay 0 (subsequence 0)
  uint32_t o0, o1, o2, o3; // cached outputs (named registers: no dynamic indexing, no stack)
  uint32_t pos;            // next output to hand out, 0..3

  __device__ __forceinline__ static void round_(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
      uint32_t k0, uint32_t k1)
  {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }

  __device__ __forceinline__ void generate()
  {
    uint32_t c0 = ctr0, c1 = ctr1, c2 = 0u, c3 = 0u, k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round_(c0, c1, c2, c3, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
  }

  __device__ __forceinline__ void bump()
  {
    if (++ctr0 == 0u)
      ++ctr1;
  }

  __device__ __forceinline__ void init(unsigned long long seed, unsigned long long offset)
  {
    key0 = (uint32_t)seed;
    key1 = (uint32_t)(seed >> 32);
    const unsigned long long blocks = offset >> 2;
    ctr0 = (uint32_t)blocks;
    ctr1 = (uint32_t)(blocks >> 32);
    pos = (uint32_t)(offset & 3ull);
    generate();
  }

  __device__ __forceinline__ uint32_t next()
  {
    const uint32_t r = pos == 0u ? o0 : (pos == 1u ? o1 : (pos == 2u ? o2 : o3));
    if (++pos == 4u) {
      bump();
      generate();
      pos = 0u;
    }
    return r;
  }

  // _curand_uniform: x * 2^-32 + 2^-33  in (0,1]  (curand_uniform.h:69-72)
  __device__ __forceinline__ static float toUniform(uint32_t x)
  {
    return __fmaf_rn((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
  }

  __device__ __forceinline__ float uniform() { return toUniform(next()); }

  // curand_uniform4 (curand4: curand_kernel.h:926-960): four consecutive outputs
  __device__ __forceinline__ float4 uniform4()
  {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      r[i] = next();
    return make_float4(toUniform(r[0]), toUniform(r[1]), toUniform(r[2]), toUniform(r[3]));
  }
};

// ---------------------------------------------------------------------------------------
// camera (gpu/cameraCreateRay.h:38-81)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float mixf(float a, float b, float t)
{
  // glm::mix for floats: x * (1 - a) + y * a
  return __fmaf_rn(b, t, __fmul_rn(a, __fsub_rn(1.0f, t)));
}

__device__ __forceinline__ void uniformSampleDisk(float radius, float rx, float ry, float &ox, float &oy)
{
  // gpu/gpu_math.h uniformSampleDisk: r = sqrt(s.x)*radius, phi = 2*pi*s.y
  const float r = __fmul_rn(__fsqrt_rn(rx), radius);
  const float phi = __fmul_rn(2.0f * 3.14159265358979323846f, ry);
  ox = __fmul_rn(r, cosf(phi));
  oy = __fmul_rn(r, sinf(phi));
}

__device__ __forceinline__ void cameraCreateRay(
    const CameraDev &c, float sx, float sy, float rz, float rw, float3 &org, float3 &dir)
{
  sx = mixf(c.region.x, c.region.z, sx);
  sy = mixf(c.region.y, c.region.w, sy);
  if (c.type == 0) {
    org = c.pos;
    dir = madd3(c.dv, sy, madd3(c.du, sx, c.p00));
    if (c.scaledAperture > 0.f) {
      float lx, ly;
      uniformSampleDisk(c.scaledAperture, rz, rw, lx, ly);
      const float3 lp = madd3(c.dv, __fmul_rn(ly, c.aspect), lx * c.du);
      org = org + lp;
      dir = dir - lp;
    }
    dir = normalize3(dir);
  } else {
    dir = c.dir;
    org = madd3(c.dv, sy, madd3(c.du, sx, c.p00));
  }
}

// getBackgroundImage, gpu/gpu_util.h:289-296: the constant colour, or a bilinear fetch of the RGBA8 background
// texture (clamp, normalised coordinates) at the pixel-sample's screen coordinate
__device__ __forceinline__ float4 backgroundAt(cudaTextureObject_t bgTex, const float4 constant, float sx, float sy)
{
  if (bgTex)
    return tex2D<float4>(bgTex, sx, sy);
  return constant;
}

// ---------------------------------------------------------------------------------------
// output encoding (gpu/gpu_util.h:323-391; glm gtc/color_space.inl:10-28, func_packing.inl:67-80)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

__device__ __forceinline__ float linearToSrgb(float c)
{
  c = clamp01(c);
  const float hi = __fmaf_rn(powf(c, 0.41666f), 1.055f, -0.055f);
  const float lo = __fmul_rn(c, 12.92f);
  return (c < 0.0031308f) ? lo : hi;
}

__device__ __forceinline__ uint32_t packUnorm4x8(float r, float g, float b, float a)
{
  const uint32_t R = (uint32_t)(unsigned char)roundf(__fmul_rn(clamp01(r), 255.0f));
  const uint32_t G = (uint32_t)(unsigned char)roundf(__fmul_rn(clamp01(g), 255.0f));
  const uint32_t B = (uint32_t)(unsigned char)roundf(__fmul_rn(clamp01(b), 255.0f));
  const uint32_t A = (uint32_t)(unsigned char)roundf(__fmul_rn(clamp01(a), 255.0f));
  return R | (G << 8) | (B << 16) | (A << 24);
}

// writeOutputColor, gpu_util.h:375-391
__device__ __forceinline__ void writeOutputColor(
    const BuffersDev &fb, int format, float4 accum, uint32_t idx, int frameIDplusOffset)
{
  const float div = float(frameIDplusOffset + 1);
  float4 c = make_float4(__fdiv_rn(accum.x, div), __fdiv_rn(accum.y, div), __fdiv_rn(accum.z, div),
      __fdiv_rn(accum.w, div));
  const float m = fmaxf(1e-12f, __fsub_rn(1.0f, fmaxf(fmaxf(c.x, c.y), c.z))); // inverseTonemap
  c.x = __fdiv_rn(c.x, m);
  c.y = __fdiv_rn(c.y, m);
  c.z = __fdiv_rn(c.z, m);
  if (format == 0) {
    fb.outF32[idx] = c;
    if (fb.outMirror)
      __stcs(reinterpret_cast<float4 *>(fb.outMirror) + idx, c);
    return;
  }
  const uint32_t v = format == 2 ? packUnorm4x8(linearToSrgb(c.x), linearToSrgb(c.y), linearToSrgb(c.z), c.w)
                                 : packUnorm4x8(c.x, c.y, c.z, c.w);
  fb.outU32[idx] = v;
  if (fb.outMirror)
    __stcs(reinterpret_cast<uint32_t *>(fb.outMirror) + idx, v);
}

// ---------------------------------------------------------------------------------------
// transfer-function lookup from a table in shared (or global) memory reproducing
// tex1D<float4> on a 256-texel clamp/normalised/linear texture (volumeIntegration.h:46-62,
// TransferFunction1D.cpp:152-186).  Texture-unit model measured on B200
// (profiles/texture_unit_model.md): xB = u*N - 0.5 in fp32, converted to 1.8 fixed point with
// round-half-up, clamped to [0, N-1]; result = ((256-k)*T[i] + k*T[i+1]) / 256.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 tfLookup(const float4 *__restrict__ tf, float coord)
{
  const float xb = __fsub_rn(__fmul_rn(coord, 256.0f), 0.5f);
  int q = __float2int_rd(__fmaf_rn(xb, 256.0f, 0.5f)); // NaN converts to 0
  q = max(0, min(q, 255 * 256));
  const int i = q >> 8;
  const float w1 = (float)(q & 255) * (1.0f / 256.0f);
  const float w0 = 1.0f - w1; // exact
  const float4 a = tf[i];
  const float4 b = tf[min(i + 1, 255)];
  return make_float4(__fmaf_rn(a.x, w0, __fmul_rn(b.x, w1)), __fmaf_rn(a.y, w0, __fmul_rn(b.y, w1)),
      __fmaf_rn(a.z, w0, __fmul_rn(b.z, w1)), __fmaf_rn(a.w, w0, __fmul_rn(b.w, w1)));
}

// position(v, range): gpu/gpu_math.h:176-180  (clamp, then multiply by the reciprocal)
__device__ __forceinline__ float rangePosition(float v, float lo, float hi)
{
  v = fmaxf(lo, fminf(v, hi));
  return __fmul_rn(__fsub_rn(v, lo), __fdiv_rn(1.0f, __fsub_rn(hi, lo)));
}

} // namespace dvr
