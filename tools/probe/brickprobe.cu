// brickprobe.cu — speed-of-light probe for a brick-staged march (VERDICT r01 "missing 5": the brick-ordered,
// shared-memory-staged field path of BASELINE.json's north_star, A/B against the texture path).
//
// What it measures, on the C2 field size (1024^3 f32): every 32x16x16-voxel brick of the volume, with the one-voxel
// apron trilinear taps need, is brought into shared memory by the TMA (cp.async.bulk.tensor.3d from the plain
// row-major volume, box 36x17x17, or cp.async.bulk of pre-gathered apron bricks), and `active` threads of the CTA
// take `spt` dependent samples each from it: software trilinear with the texture unit's measured 1.8 fixed-point
// weight rule (profiles/texture_unit_model.md), transfer-function lookup from shared memory, opacity correction
// (powf) and the front-to-back composite of gpu/volumeIntegration.h:64-103 — the inner loop a brick march would run.
// No ray queues, no per-ray set-up, no ordering between bricks, no frame buffer: whatever a real brick-staged march
// adds comes ON TOP of these times, so the numbers are an upper bound on its frame rate.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o brickprobe brickprobe.cu
//   ./brickprobe [n=1024]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                        \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) {                                                                         \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      std::exit(1);                                                                                  \
    }                                                                                                \
  } while (0)

#ifndef PBY
#define PBY 16
#endif
#ifndef PBZ
#define PBZ 16
#endif
constexpr int BX = 32, BY = PBY, BZ = PBZ;        // brick core (voxels whose lower tap the brick owns)
constexpr int SX = 36, SY = BY + 1, SZ = BZ + 1;  // staged box: x padded to a multiple of 16 bytes
constexpr int BOX = SX * SY * SZ;                 // floats per staged brick (32x16x16: 10404 -> 41616 B)
constexpr int STAGE = (BOX * 4 + 127) / 128 * 32; // floats between two stages: TMA destinations are 128-byte aligned
constexpr int THREADS = 128;

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void barInit(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void barExpectTx(uint64_t *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
// bounded wait: a TMA that never lands must not hang the GPU
__device__ __forceinline__ bool barWait(uint64_t *bar, unsigned phase)
{
  const uint32_t a = smemAddr(bar);
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(phase) : "memory");
    if (ok)
      return true;
  }
  return false;
}
__device__ __forceinline__ void tmaLoad3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smemAddr(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void bulkLoad1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// tex1D<float4> on the 256-texel transfer function, measured unit arithmetic (csrc/dvr_device.cuh::tfLookup)
__device__ __forceinline__ float4 tfLookup(const float4 *tf, float coord)
{
  const float xb = __fsub_rn(__fmul_rn(coord, 256.0f), 0.5f);
  int q = __float2int_rd(__fmaf_rn(xb, 256.0f, 0.5f));
  q = max(0, min(q, 255 * 256));
  const int i = q >> 8;
  const float w1 = (float)(q & 255) * (1.0f / 256.0f), w0 = 1.0f - w1;
  const float4 a = tf[i], b = tf[min(i + 1, 255)];
  return make_float4(__fmaf_rn(a.x, w0, __fmul_rn(b.x, w1)), __fmaf_rn(a.y, w0, __fmul_rn(b.y, w1)),
      __fmaf_rn(a.z, w0, __fmul_rn(b.z, w1)), __fmaf_rn(a.w, w0, __fmul_rn(b.w, w1)));
}

__device__ __forceinline__ float wf(int w) { return __int_as_float(0x4B000000 | w) - 8388608.0f; } // exact int -> float

// tex3D<float> with linear filtering from a staged brick: x/y/z are brick-local unnormalised coordinates (xB of the
// model) of a sample whose lower taps lie in the brick core.  Weights: the unit's integer rule, 1/256 units.
__device__ __forceinline__ float brickTrilinear(const float *__restrict__ b, float x, float y, float z)
{
  const int qx = __float2int_rd(__fmaf_rn(x, 256.0f, 0.5f)), qy = __float2int_rd(__fmaf_rn(y, 256.0f, 0.5f)),
            qz = __float2int_rd(__fmaf_rn(z, 256.0f, 0.5f));
  const int ix = qx >> 8, iy = qy >> 8, iz = qz >> 8, kx = qx & 255, ky = qy & 255, kz = qz & 255;
  const float *p = b + (iz * SY + iy) * SX + ix;
  const float v000 = p[0], v100 = p[1], v010 = p[SX], v110 = p[SX + 1];
  const float v001 = p[SY * SX], v101 = p[SY * SX + 1], v011 = p[SY * SX + SX], v111 = p[SY * SX + SX + 1];
  float acc = 0.f;
  {
    const int B = 256 - kz;
    const int X1 = (B * kx + 128) >> 8, X0 = B - X1;
    const int w11 = (X1 * ky + 128) >> 8, w10 = X1 - w11, w01 = (X0 * ky + 127) >> 8, w00 = X0 - w01;
    acc = __fmaf_rn(v000, wf(w00), acc);
    acc = __fmaf_rn(v100, wf(w10), acc);
    acc = __fmaf_rn(v010, wf(w01), acc);
    acc = __fmaf_rn(v110, wf(w11), acc);
  }
  {
    const int B = kz;
    const int X1 = (B * kx + 128) >> 8, X0 = B - X1;
    const int w11 = (X1 * ky + 128) >> 8, w10 = X1 - w11, w01 = (X0 * ky + 127) >> 8, w00 = X0 - w01;
    acc = __fmaf_rn(v001, wf(w00), acc);
    acc = __fmaf_rn(v101, wf(w10), acc);
    acc = __fmaf_rn(v011, wf(w01), acc);
    acc = __fmaf_rn(v111, wf(w11), acc);
  }
  return acc * (1.0f / 256.0f);
}

struct ProbeArgs
{
  const float *bricks1d; // mode 1: pre-gathered apron bricks, BOX floats each, brick order x-fastest
  const float4 *tf;
  float4 *out;
  unsigned int *counter;
  int nbx, nby, nbz;
  int mode;   // 0: TMA 3-D box from the row-major volume, 1: 1-D bulk copy of pre-gathered bricks
  int active; // threads per CTA that take samples
  int spt;    // dependent samples per active thread and brick
  float exponent;
  unsigned int *timeouts;
};

template <int ST>
__global__ void __launch_bounds__(THREADS) brickProbeKernel(const __grid_constant__ CUtensorMap map, const ProbeArgs A)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  float *buf = reinterpret_cast<float *>(smemRaw);
  __shared__ float4 s_tf[256];
  __shared__ uint64_t bar[ST];
  __shared__ int ids[ST];
  const int tid = threadIdx.x;
  const int nBricks = A.nbx * A.nby * A.nbz;
  for (int i = tid; i < 256; i += THREADS)
    s_tf[i] = A.tf[i];
  if (tid == 0) {
    for (int s = 0; s < ST; ++s)
      barInit(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s) {
    const int id = (int)atomicAdd(A.counter, 1u);
    ids[s] = id;
    if (id < nBricks) {
      barExpectTx(&bar[s], BOX * 4);
      if (A.mode == 0) {
        const int bx = id % A.nbx, by = (id / A.nbx) % A.nby, bz = id / (A.nbx * A.nby);
        tmaLoad3d(buf + (size_t)s * STAGE, &map, bx * BX, by * BY, bz * BZ, &bar[s]);
      } else {
        bulkLoad1d(buf + (size_t)s * STAGE, A.bricks1d + (size_t)id * BOX, BOX * 4, &bar[s]);
      }
    }
  };
  if (tid == 0)
    for (int s = 0; s < ST; ++s)
      issue(s);
  __syncthreads();

  // a "ray" per active thread: direction from the thread id, fixed for the run
  const float ang = 0.37f * (float)tid;
  float3 d = make_float3(0.47f + 0.05f * cosf(ang), 0.34f + 0.05f * sinf(ang), 0.81f);
  float3 color = make_float3(0.f, 0.f, 0.f);
  float opacity = 0.f, transmittance = 1.f;
  float3 p = make_float3((float)(tid % 8) * 3.8f + 0.3f, (float)((tid / 8) % 4) * 3.8f + 0.2f, 0.1f * (float)(tid & 3));

  for (int it = 0;; ++it) {
    const int s = it % ST;
    const unsigned phase = (unsigned)(it / ST) & 1u;
    const int id = ids[s];
    if (id >= nBricks)
      break;
    if (!barWait(&bar[s], phase)) {
      if (tid == 0)
        atomicAdd(A.timeouts, 1u);
      break;
    }
    const float *b = buf + (size_t)s * STAGE;
    if (tid < A.active) {
      for (int k = 0; k < A.spt; ++k) {
        // next lattice point of this ray, wrapped back into the brick core
        p.x += d.x; p.y += d.y; p.z += d.z;
        if (p.x >= (float)BX - 0.01f) p.x -= (float)BX - 0.02f;
        if (p.y >= (float)BY - 0.01f) p.y -= (float)BY - 0.02f;
        if (p.z >= (float)BZ - 0.01f) p.z -= (float)BZ - 0.02f;
        const float v = brickTrilinear(b, p.x, p.y, p.z);
        const float c = fmaxf(0.f, fminf(v, 1.f));
        const float4 co = tfLookup(s_tf, c);
        const float st = powf(__fsub_rn(1.f, co.w), A.exponent);
        if (opacity < 0.99f) {
          const float w = __fmul_rn(transmittance, __fsub_rn(1.f, st));
          color.x = __fmaf_rn(w, co.x, color.x);
          color.y = __fmaf_rn(w, co.y, color.y);
          color.z = __fmaf_rn(w, co.z, color.z);
          opacity = __fadd_rn(opacity, w);
          transmittance = __fmul_rn(transmittance, st);
        }
      }
      if (opacity >= 0.99f) { // keep the loop doing the full work for the whole run
        opacity = 0.f;
        transmittance = 1.f;
      }
    }
    __syncthreads(); // every reader is done with stage s
    if (tid == 0)
      issue(s);
    __syncthreads();
  }
  if (tid < A.active)
    A.out[(size_t)blockIdx.x * THREADS + tid] = make_float4(color.x, color.y, color.z, opacity + transmittance);
}

__global__ void fillKernel(float *v, size_t n, int dim)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % dim), y = (int)((i / dim) % dim), z = (int)(i / ((size_t)dim * dim));
    v[i] = 0.5f + 0.25f * __sinf(0.02f * (float)x) * __cosf(0.03f * (float)y) + 0.2f * __sinf(0.011f * (float)z);
  }
}

__global__ void gatherKernel(const float *vol, float *bricks, int dim, int nbx, int nby, int nbz)
{
  const size_t n = (size_t)nbx * nby * nbz * BOX;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t id = i / BOX;
    const int r = (int)(i % BOX);
    const int lx = r % SX, ly = (r / SX) % SY, lz = r / (SX * SY);
    const int bx = (int)(id % nbx), by = (int)((id / nbx) % nby), bz = (int)(id / ((size_t)nbx * nby));
    const int x = min(bx * BX + lx, dim - 1), y = min(by * BY + ly, dim - 1), z = min(bz * BZ + lz, dim - 1);
    bricks[i] = vol[((size_t)z * dim + y) * dim + x];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ST>
static float runConfig(const CUtensorMap &map, ProbeArgs A, int ctasPerSm, int sms, int reps)
{
  const size_t smem = (size_t)ST * STAGE * 4;
  CK(cudaFuncSetAttribute(brickProbeKernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, brickProbeKernel<ST>, THREADS, smem));
  if (ctasPerSm > occ)
    ctasPerSm = occ;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaMemsetAsync(A.counter, 0, 4));
    CK(cudaEventRecord(e0));
    brickProbeKernel<ST><<<sms * ctasPerSm, THREADS, smem>>>(map, A);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2) {
      sum += ms;
      best = ms < best ? ms : best;
    }
  }
  std::printf("  stages %d  CTAs/SM %d (max %d)  active %3d  spt %3d  mode %s : %.4f ms avg, %.4f ms best\n", ST, ctasPerSm, occ,
      A.active, A.spt, A.mode == 0 ? "tma3d " : "bulk1d", sum / reps, best);
  return sum / reps;
}

int main(int argc, char **argv)
{
  const int dim = argc > 1 ? std::atoi(argv[1]) : 1024;
  const bool quick = argc > 2;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t nvox = (size_t)dim * dim * dim;
  float *vol = nullptr;
  CK(cudaMalloc(&vol, nvox * 4));
  fillKernel<<<sms * 8, 256>>>(vol, nvox, dim);
  CK(cudaDeviceSynchronize());
  const int nbx = dim / BX, nby = dim / BY, nbz = dim / BZ;
  const size_t nBricks = (size_t)nbx * nby * nbz;
  float *bricks = nullptr;
  CK(cudaMalloc(&bricks, nBricks * BOX * 4));
  gatherKernel<<<sms * 8, 256>>>(vol, bricks, dim, nbx, nby, nbz);
  CK(cudaDeviceSynchronize());

  // transfer function: grey ramp, alpha rising to 0.6 (most rays of C2 never saturate)
  std::vector<float> tf(256 * 4);
  for (int i = 0; i < 256; ++i) {
    tf[4 * i] = tf[4 * i + 1] = tf[4 * i + 2] = i / 255.f;
    tf[4 * i + 3] = 0.6f * i / 255.f;
  }
  float4 *dtf = nullptr;
  CK(cudaMalloc(&dtf, 256 * 16));
  CK(cudaMemcpy(dtf, tf.data(), 256 * 16, cudaMemcpyHostToDevice));
  float4 *out = nullptr;
  CK(cudaMalloc(&out, (size_t)sms * 8 * THREADS * 16));
  unsigned int *counter = nullptr, *timeouts = nullptr;
  CK(cudaMalloc(&counter, 4));
  CK(cudaMalloc(&timeouts, 4));
  CK(cudaMemset(timeouts, 0, 4));

  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) {
    std::fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  CUtensorMap map;
  std::memset(&map, 0, sizeof(map));
  const cuuint64_t gdim[3] = {(cuuint64_t)dim, (cuuint64_t)dim, (cuuint64_t)dim};
  const cuuint64_t gstride[2] = {(cuuint64_t)dim * 4, (cuuint64_t)dim * dim * 4};
  const cuuint32_t box[3] = {SX, SY, SZ};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vol, gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    std::fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)cr);
    return 1;
  }

  const double volGB = (double)nvox * 4 / 1e9, boxGB = (double)nBricks * BOX * 4 / 1e9;
  std::printf("brickprobe: %d^3 f32 (%.3f GB), %zu bricks of %dx%dx%d (+apron -> %dx%dx%d, %.3f GB staged), %d SMs\n", dim,
      volGB, nBricks, BX, BY, BZ, SX, SY, SZ, boxGB, sms);
  ProbeArgs A{bricks, dtf, out, counter, nbx, nby, nbz, 0, 0, 0, 0.5f / 256.f * 256.f / 256.f, timeouts};
  A.exponent = 1.0f / 256.0f;

  auto report = [&](const char *what, float ms, double samples) {
    std::printf("    -> %s: %.3f ms = %.0f frames/s bound; volume bytes / time = %.0f GB/s, staged bytes / time = %.0f GB/s",
        what, ms, 1000.0 / ms, volGB / ms * 1e3, boxGB / ms * 1e3);
    if (samples > 0)
      std::printf(", %.1f Gsamples/s", samples / ms / 1e6);
    std::printf("\n");
  };

  // 1. pure streaming ceiling (no samples)
  std::printf("streaming only:\n");
  for (int mode = 0; mode < 2; ++mode) {
    A.mode = mode; A.active = 0; A.spt = 0;
    report("stream", runConfig<1>(map, A, 5, sms, 5), 0);
    if (!quick) {
      report("stream", runConfig<1>(map, A, 3, sms, 5), 0);
      report("stream", runConfig<2>(map, A, 2, sms, 5), 0);
    }
  }
  // 2. C2 at 1 voxel per step: 76.05 M samples per frame = 580 per brick; at 0.5 voxel per step: 1160 per brick
  const int cfg[][2] = {{48, 12}, {96, 6}, {128, 5}, {48, 24}, {96, 12}, {128, 9}};
  for (int mode = 0; mode < 2; ++mode) {
    std::printf("streaming + sampling, mode %s:\n", mode == 0 ? "tma3d" : "bulk1d");
    for (const auto &c : cfg) {
      A.mode = mode; A.active = c[0]; A.spt = c[1];
      const double samples = (double)nBricks * c[0] * c[1];
      report("1 stage x5", runConfig<1>(map, A, 5, sms, 5), samples);
      if (!quick)
        report("2 stages x2", runConfig<2>(map, A, 2, sms, 5), samples);
    }
  }
  unsigned int to = 0;
  CK(cudaMemcpy(&to, timeouts, 4, cudaMemcpyDeviceToHost));
  std::printf("barrier wait timeouts: %u\n", to);
  return to ? 2 : 0;
}
