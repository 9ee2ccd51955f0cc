// dvr_scene.cu — the mixed-scene frame kernel (surfaces + volumes + lights) and the host side of
// DvrSurfaces (geometry upload, BVH build).  Target: sm_100a only.
//
// Kernel = the raygen programs of renderer/DirectLight_ptx.cu:294-418 and renderer/Raycast_ptx.cu:60-179 with their
// surface branch: closest surface, volumes marched up to surfaceHit.t, matte shading with shadow rays through
// surfaces and volumes, and the loop behind translucent surfaces.  Pixels whose rays meet no surface run exactly the
// arithmetic of dvrFrameKernel (same marcher, same Philox stream), so they are bit-identical to dvr_render.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "dvr_internal.h"
#include "dvr_frame_common.cuh"
#include "dvr_scene.cuh"

namespace dvr {

constexpr int kMaxInlineLights = 16;

struct SceneLaunch
{
  FrameLaunch f;
  SceneDev sc;
  LightDev lights[kMaxInlineLights];
};

template <bool SKIP>
__global__ void __launch_bounds__(kBlockThreads, 2) dvrSceneFrameKernel(const __grid_constant__ SceneLaunch S)
{
  __shared__ float4 s_tf[kMaxInlineInstances * DVR_TF_SIZE];
  const FrameLaunch &P = S.f;
  SceneDev sc = S.sc;
  sc.lights = S.lights;

  const int lane = threadIdx.x & 31;
  const InstanceDev *inst = (P.nInst <= kMaxInlineInstances) ? P.inl : P.ext;
  const int nInst = P.nInst;
  {
    const int nTab = min(P.nInst, kMaxInlineInstances);
    for (int i = threadIdx.x; i < nTab * DVR_TF_SIZE; i += blockDim.x)
      s_tf[i] = __ldg(&inst[i / DVR_TF_SIZE].v.tf[i % DVR_TF_SIZE]);
    __syncthreads();
  }
  const TfSelectShared tfOf{s_tf, inst};

  MarchStats st{0ull, 0ull};
  const uint32_t nTiles = P.tilesW * P.tilesH;
  const bool centered = P.integrator == DVR_INTEGRATOR_RAYCAST;
  const bool initFrame = P.frameID == 0 && P.checkerboardID <= 0;
  const AccumCtx actx{P.width, P.height, P.format, P.frameID, P.checkerboardID, P.fb};
  const int iterations = centered ? 1 : P.numIterations; // the raycast raygen has no sample loop

  for (uint32_t tile = nextTile(P.sched, lane); tile < nTiles; tile = nextTile(P.sched, lane)) {
    const uint32_t tyIdx = P.tileY0 + tile / P.tilesW, txIdx = P.tileX0 + tile % P.tilesW;
    if (P.tileRanks > 1u && ((tyIdx / P.tileBand) % P.tileRanks) != P.tileRank)
      continue;
    const uint32_t lx = txIdx * kTileW + (lane % kTileW), ly = tyIdx * kTileH + (lane / kTileW);
    if (lx >= P.launchW || ly >= P.launchH)
      continue;
    uint32_t px = lx, py = ly;
    if (P.checkerboardID >= 0) { // createScreenSample.h:38-46
      px = lx * 2u + (uint32_t)(P.checkerboardID & 1);
      py = ly * 2u + (uint32_t)((P.checkerboardID >> 1) & 1);
    }
    if (px >= P.width || py >= P.height)
      continue;

    Philox rng;
    rng.init((unsigned long long)(int)(py * P.width + px), (unsigned long long)P.frameID * 512ull);

    for (int it = 0; it < iterations; ++it) {
      const float4 r = rng.uniform4(); // makePrimaryRay, cameraCreateRay.h:74-81
      const float sx = __fmul_rn(centered ? (float)px : __fadd_rn((float)px, r.x), P.invW);
      const float sy = __fmul_rn(centered ? (float)py : __fadd_rn((float)py, r.y), P.invH);
      float3 org, dir;
      cameraCreateRay(P.cam, sx, sy, r.z, r.w, org, dir);

      float3 outputColor = f3(0.f, 0.f, 0.f);
      float3 outputNormal = dir;
      float outputOpacity = 0.f;
      float depth = 1e30f;
      uint32_t primID = ~0u, objID = ~0u, instID = ~0u;
      bool firstHit = true;
      float tLower = 0.f; // ray.t.lower

      while (outputOpacity < 0.99f) {
        SurfaceHitDev hit;
        intersectSurfaceClosest(sc, org, dir, tLower, FLT_MAX, hit);
        float3 color = f3(0.f, 0.f, 0.f);
        float opacity = 0.f;
        uint32_t vObjID = ~0u, vInstID = ~0u;
        bool anyHit = false;
        if (hit.found) {
          const float vDepth = rayMarchAllVolumes<SKIP, false, false, false, -1>(inst, nInst, tfOf, org, dir, hit.t,
              P.invSamplingRate, rng, color, opacity, vObjID, vInstID, st, nullptr, anyHit, tLower);
          if (firstHit) {
            if (vDepth < hit.t) { // volumeFirst
              outputNormal = f3(-dir.x, -dir.y, -dir.z);
              depth = vDepth;
              primID = 0u;
              objID = vObjID;
              instID = vInstID;
            } else {
              outputNormal = centered ? hit.Ng : hit.Ns; // Raycast_ptx.cu:116 / DirectLight_ptx.cu:350
              depth = hit.t;
              primID = hit.primID;
              objID = hit.objID;
              instID = hit.instID;
            }
            firstHit = false;
          }
          if (centered) {
            // Raycast_ptx.cu:125-135: headlight |dir . Ns| * ambientColor on the material tint, blended behind the
            // volume segment in front of the surface
            const float lit = fabsf(dot3(dir, hit.Ns));
            const SceneSurfaceDev &sd = *hit.surface;
            const float om = __fsub_rn(1.f, opacity);
            color.x = __fmaf_rn(sd.baseColor.x * (lit * sc.ambientColor.x), om, color.x);
            color.y = __fmaf_rn(sd.baseColor.y * (lit * sc.ambientColor.y), om, color.y);
            color.z = __fmaf_rn(sd.baseColor.z * (lit * sc.ambientColor.z), om, color.z);
            opacity = __fmaf_rn(sd.opacity, om, opacity);
          } else {
            // DirectLight_ptx.cu:359-368.  Reference quirk (kept for parity): the shading result REPLACES the colour and
            // opacity of the volume segment marched in front of the surface, and is then blended onto itself.
            const float4 sh = shadeSurfaceDirectLight<SKIP>(sc, inst, nInst, tfOf, dir, hit, P.invSamplingRate, rng);
            if (isnan(sh.x) || isnan(sh.y) || isnan(sh.z)) {
              color = f3(0.f, 0.f, 0.f);
              opacity = 0.f;
            } else {
              color = f3(sh.x, sh.y, sh.z);
              opacity = sh.w;
            }
            const float om = __fsub_rn(1.f, opacity);
            color.x = __fmaf_rn(sh.x, om, color.x);
            color.y = __fmaf_rn(sh.y, om, color.y);
            color.z = __fmaf_rn(sh.z, om, color.z);
            opacity = __fmaf_rn(sh.w, om, opacity);
          }
          color = color * opacity;
          const float oo = __fsub_rn(1.f, outputOpacity);
          outputColor.x = __fmaf_rn(color.x, oo, outputColor.x);
          outputColor.y = __fmaf_rn(color.y, oo, outputColor.y);
          outputColor.z = __fmaf_rn(color.z, oo, outputColor.z);
          outputOpacity = __fmaf_rn(opacity, oo, outputOpacity);
          tLower = __fadd_rn(hit.t, hit.epsilon);
        } else {
          const float volumeDepth = rayMarchAllVolumes<SKIP, false, false, false, -1>(inst, nInst, tfOf, org, dir,
              FLT_MAX, P.invSamplingRate, rng, color, opacity, vObjID, vInstID, st, nullptr, anyHit, tLower);
          if (firstHit) {
            depth = fminf(depth, volumeDepth);
            primID = 0u;
            objID = vObjID;
            instID = vInstID;
          }
          color = color * opacity;
          const float4 bg = backgroundAt(P.bgTex, P.background, sx, sy);
          const float om = __fsub_rn(1.f, opacity);
          color.x = __fmaf_rn(bg.x, om, color.x);
          color.y = __fmaf_rn(bg.y, om, color.y);
          color.z = __fmaf_rn(bg.z, om, color.z);
          opacity = __fmaf_rn(bg.w, om, opacity);
          const float oo = __fsub_rn(1.f, outputOpacity);
          outputColor.x = __fmaf_rn(color.x, oo, outputColor.x);
          outputColor.y = __fmaf_rn(color.y, oo, outputColor.y);
          outputColor.z = __fmaf_rn(color.z, oo, outputColor.z);
          outputOpacity = __fmaf_rn(opacity, oo, outputOpacity);
          break;
        }
      }
      accumResults(actx, px, py, make_float4(outputColor.x, outputColor.y, outputColor.z, outputOpacity), depth,
          outputColor, outputNormal, primID, objID, instID, it, initFrame && it == 0);
    }
  }
  retireWarp(P.sched, lane);
}

template <bool SKIP>
static int launchSceneT(const SceneLaunch &L, cudaStream_t s)
{
  static int blocksPerSm = 0;
  if (blocksPerSm == 0) {
    DVR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, dvrSceneFrameKernel<SKIP>, kBlockThreads, 0));
    if (blocksPerSm < 1)
      blocksPerSm = 1;
  }
  const uint32_t nTiles = L.f.tilesW * L.f.tilesH;
  const uint32_t warpsPerBlock = kBlockThreads / 32;
  uint32_t grid = (uint32_t)(smCount() * blocksPerSm);
  const uint32_t need = (nTiles + warpsPerBlock - 1) / warpsPerBlock;
  if (grid > need)
    grid = need;
  if (grid == 0)
    grid = 1;
  dvrSceneFrameKernel<SKIP><<<grid, kBlockThreads, 0, s>>>(L);
  DVR_CUDA(cudaGetLastError());
  countLaunch();
  return DVR_OK;
}

} // namespace dvr

// ---------------------------------------------------------------------------------------------------------------
// host side: DvrSurfaces
// ---------------------------------------------------------------------------------------------------------------
using namespace dvr;

struct DvrSurfaces
{
  std::vector<void *> allocations;
  BvhNode *nodes = nullptr;
  ScenePrimRef *prims = nullptr;
  SceneSurfaceDev *surfaces = nullptr;
  uint32_t nNodes = 0, nPrims = 0, nSurfaces = 0;
  int device = 0;
};

namespace {

struct BuildPrim
{
  float lo[3], hi[3], c[3];
  ScenePrimRef ref;
};

void invert3x4(const float *m, float *out)
{
  // rows of m: (a b c | t); inverse of the 3x3 in double, then -inv * t
  const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  const double id = det != 0.0 ? 1.0 / det : 0.0;
  const double r[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id,
      (a * i - c * g) * id, (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
  for (int row = 0; row < 3; ++row) {
    out[4 * row + 0] = (float)r[3 * row + 0];
    out[4 * row + 1] = (float)r[3 * row + 1];
    out[4 * row + 2] = (float)r[3 * row + 2];
    out[4 * row + 3] = (float)-(r[3 * row + 0] * m[3] + r[3 * row + 1] * m[7] + r[3 * row + 2] * m[11]);
  }
}

void buildNode(std::vector<BvhNode> &nodes, std::vector<BuildPrim> &prims, uint32_t nodeIdx, uint32_t first,
    uint32_t count)
{
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t k = first; k < first + count; ++k)
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], prims[k].lo[a]);
      hi[a] = std::max(hi[a], prims[k].hi[a]);
      clo[a] = std::min(clo[a], prims[k].c[a]);
      chi[a] = std::max(chi[a], prims[k].c[a]);
    }
  BvhNode n;
  n.lo = make_float3(lo[0], lo[1], lo[2]);
  n.hi = make_float3(hi[0], hi[1], hi[2]);
  int axis = 0;
  for (int a = 1; a < 3; ++a)
    if (chi[a] - clo[a] > chi[axis] - clo[axis])
      axis = a;
  if (count <= 4u || !(chi[axis] - clo[axis] > 0.f)) {
    n.leftOrFirst = first;
    n.count = count;
    nodes[nodeIdx] = n;
    return;
  }
  const uint32_t mid = first + count / 2u;
  std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count,
      [axis](const BuildPrim &x, const BuildPrim &y) { return x.c[axis] < y.c[axis]; });
  const uint32_t left = (uint32_t)nodes.size();
  nodes.emplace_back();
  nodes.emplace_back();
  n.leftOrFirst = left;
  n.count = 0u;
  nodes[nodeIdx] = n;
  buildNode(nodes, prims, left, first, mid - first);
  buildNode(nodes, prims, left + 1u, mid, first + count - mid);
}

template <typename T>
int upload(DvrSurfaces *S, const T *host, size_t n, const T **dev)
{
  *dev = nullptr;
  if (!host || n == 0)
    return DVR_OK;
  void *p = nullptr;
  DVR_CUDA(cudaMalloc(&p, n * sizeof(T)));
  S->allocations.push_back(p);
  DVR_CUDA(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
  *dev = (const T *)p;
  return DVR_OK;
}

} // namespace

extern "C" {

int dvr_surfaces_destroy(DvrSurfaces *s)
{
  if (!s)
    return DVR_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(s->device);
  for (void *p : s->allocations)
    cudaFree(p);
  cudaSetDevice(prev);
  delete s;
  return DVR_OK;
}

int dvr_surfaces_create(const DvrSurfaceDesc *surfaces, uint32_t nSurfaces, void *stream, DvrSurfaces **out)
{
  (void)stream;
  if (!out || (nSurfaces && !surfaces)) {
    setError("dvr_surfaces_create: null argument");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  if (dvr_device_count() <= 0) {
    setError("dvr_surfaces_create: no CUDA device (this library has no CPU fallback)");
    return DVR_ERR_NO_DEVICE;
  }
  auto *S = new DvrSurfaces();
  cudaGetDevice(&S->device);
  std::vector<SceneSurfaceDev> devSurf(nSurfaces);
  std::vector<BuildPrim> prims;
  for (uint32_t si = 0; si < nSurfaces; ++si) {
    const DvrSurfaceDesc &d = surfaces[si];
    const bool tri = d.geometryType == DVR_GEOMETRY_TRIANGLE;
    if ((!tri && d.geometryType != DVR_GEOMETRY_SPHERE) || !d.vertexPosition || d.nVertices == 0) {
      dvr_surfaces_destroy(S);
      setError("dvr_surfaces_create: surface without vertex positions or of an unknown geometry type");
      return DVR_ERR_INVALID_ARGUMENT;
    }
    const uint32_t nPrims = d.index ? d.nPrimitives : (tri ? d.nVertices / 3u : d.nVertices);
    if (d.index) // every index must address a vertex
      for (size_t k = 0; k < (size_t)nPrims * (tri ? 3u : 1u); ++k)
        if (d.index[k] >= d.nVertices) {
          dvr_surfaces_destroy(S);
          setError("dvr_surfaces_create: primitive.index addresses a vertex beyond vertex.position");
          return DVR_ERR_INVALID_ARGUMENT;
        }
    SceneSurfaceDev &sd = devSurf[si];
    std::memset(&sd, 0, sizeof(sd));
    sd.geometryType = d.geometryType;
    sd.cullBackfaces = d.cullBackfaces;
    int rc = upload(S, d.vertexPosition, (size_t)d.nVertices * 3, &sd.vertices);
    if (rc == DVR_OK)
      rc = upload(S, d.index, (size_t)nPrims * (tri ? 3u : 1u), &sd.index);
    if (rc == DVR_OK && tri)
      rc = upload(S, d.vertexNormal, (size_t)d.nVertices * 3, &sd.normals);
    if (rc == DVR_OK && !tri)
      rc = upload(S, d.vertexRadius, (size_t)d.nVertices, &sd.radii);
    if (rc == DVR_OK)
      rc = upload(S, d.primitiveId, (size_t)nPrims, &sd.primitiveId);
    if (rc != DVR_OK) {
      dvr_surfaces_destroy(S);
      return rc;
    }
    sd.radius = d.radius;
    sd.baseColor = make_float3(d.color[0], d.color[1], d.color[2]);
    { // adjustedMaterialOpacity(color.w * opacity, alphaMode, cutoff), MatteShader_ptx.cu:44-49
      const float o = d.color[3] * d.opacity;
      sd.opacity = d.alphaMode == DVR_ALPHA_OPAQUE ? 1.f : (d.alphaMode == DVR_ALPHA_BLEND ? o : (o < d.alphaCutoff ? 0.f : 1.f));
    }
    sd.surfaceId = d.surfaceId;
    sd.instanceId = d.instanceId;
    std::memcpy(sd.o2w, d.objectToWorld, sizeof(sd.o2w));
    invert3x4(sd.o2w, sd.w2o);
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    sd.identity = std::memcmp(sd.o2w, ident, sizeof(ident)) == 0;

    for (uint32_t p = 0; p < nPrims; ++p) {
      float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      if (tri) {
        for (int c = 0; c < 3; ++c) {
          const uint32_t vi = d.index ? d.index[3u * p + c] : 3u * p + c;
          for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], d.vertexPosition[3 * (size_t)vi + a]);
            hi[a] = std::max(hi[a], d.vertexPosition[3 * (size_t)vi + a]);
          }
        }
      } else {
        const uint32_t vi = d.index ? d.index[p] : p;
        const float r = std::fabs(d.vertexRadius ? d.vertexRadius[vi] : d.radius);
        for (int a = 0; a < 3; ++a) {
          lo[a] = d.vertexPosition[3 * (size_t)vi + a] - r;
          hi[a] = d.vertexPosition[3 * (size_t)vi + a] + r;
        }
      }
      BuildPrim bp;
      bp.ref = ScenePrimRef{si, p};
      for (int a = 0; a < 3; ++a) {
        bp.lo[a] = FLT_MAX;
        bp.hi[a] = -FLT_MAX;
      }
      for (int corner = 0; corner < 8; ++corner) { // world box of the object-space box
        const float x = (corner & 1) ? hi[0] : lo[0], y = (corner & 2) ? hi[1] : lo[1], z = (corner & 4) ? hi[2] : lo[2];
        for (int a = 0; a < 3; ++a) {
          const float *m = sd.o2w + 4 * a;
          const float w = m[0] * x + m[1] * y + m[2] * z + m[3];
          bp.lo[a] = std::min(bp.lo[a], w);
          bp.hi[a] = std::max(bp.hi[a], w);
        }
      }
      bool finite = true;
      for (int a = 0; a < 3; ++a) {
        const float pad = 1e-6f * std::max(std::fabs(bp.lo[a]), std::fabs(bp.hi[a])) + 1e-7f; // conservative boxes
        bp.lo[a] -= pad;
        bp.hi[a] += pad;
        bp.c[a] = 0.5f * (bp.lo[a] + bp.hi[a]);
        finite = finite && std::isfinite(bp.lo[a]) && std::isfinite(bp.hi[a]);
      }
      if (finite) // primitives with NaN / inf coordinates can never be hit
        prims.push_back(bp);
    }
  }
  std::vector<BvhNode> nodes;
  if (!prims.empty()) {
    nodes.reserve(prims.size());
    nodes.emplace_back();
    buildNode(nodes, prims, 0u, 0u, (uint32_t)prims.size());
  }
  std::vector<ScenePrimRef> refs(prims.size());
  for (size_t k = 0; k < prims.size(); ++k)
    refs[k] = prims[k].ref;
  const BvhNode *dn = nullptr;
  const ScenePrimRef *dp = nullptr;
  const SceneSurfaceDev *ds = nullptr;
  int rc = upload(S, nodes.data(), nodes.size(), &dn);
  if (rc == DVR_OK)
    rc = upload(S, refs.data(), refs.size(), &dp);
  if (rc == DVR_OK)
    rc = upload(S, devSurf.data(), devSurf.size(), &ds);
  if (rc != DVR_OK) {
    dvr_surfaces_destroy(S);
    return rc;
  }
  S->nodes = const_cast<BvhNode *>(dn);
  S->prims = const_cast<ScenePrimRef *>(dp);
  S->surfaces = const_cast<SceneSurfaceDev *>(ds);
  S->nNodes = (uint32_t)nodes.size();
  S->nPrims = (uint32_t)refs.size();
  S->nSurfaces = nSurfaces;
  *out = S;
  return DVR_OK;
}

int dvr_surfaces_info(const DvrSurfaces *s, uint32_t *nPrimitives, uint32_t *nNodes)
{
  if (!s) {
    setError("dvr_surfaces_info: null surfaces");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  if (nPrimitives)
    *nPrimitives = s->nPrims;
  if (nNodes)
    *nNodes = s->nNodes;
  return DVR_OK;
}

} // extern "C"

namespace dvr {

bool sceneHasSurfaces(const DvrSceneParams *scene) { return scene && scene->surfaces && scene->surfaces->nPrims > 0u; }

int launchSceneFrame(const FrameLaunch &f, const DvrSceneParams *scene, bool skip, cudaStream_t s)
{
  if (scene->nLights > (uint32_t)kMaxInlineLights) {
    setError("dvr_render_scene: more than 16 lights");
    return DVR_ERR_UNSUPPORTED;
  }
  if (scene->nLights && !scene->lights) {
    setError("dvr_render_scene: null light array");
    return DVR_ERR_INVALID_ARGUMENT;
  }
  SceneLaunch L;
  std::memset(&L, 0, sizeof(L));
  L.f = f;
  const DvrSurfaces *S = scene->surfaces;
  L.sc.nodes = S->nodes;
  L.sc.prims = S->prims;
  L.sc.surfaces = S->surfaces;
  L.sc.nNodes = S->nNodes;
  L.sc.nPrims = S->nPrims;
  L.sc.nLights = (int)scene->nLights;
  for (uint32_t i = 0; i < scene->nLights; ++i) {
    const DvrLight &l = scene->lights[i];
    if (l.type != DVR_LIGHT_DIRECTIONAL && l.type != DVR_LIGHT_POINT) {
      setError("dvr_render_scene: unknown light type");
      return DVR_ERR_INVALID_ARGUMENT;
    }
    L.lights[i].type = l.type;
    L.lights[i].color = make_float3(l.color[0], l.color[1], l.color[2]);
    L.lights[i].vec = make_float3(l.vec[0], l.vec[1], l.vec[2]);
    L.lights[i].strength = l.strength;
  }
  L.sc.ambientColor = make_float3(scene->ambientColor[0], scene->ambientColor[1], scene->ambientColor[2]);
  L.sc.ambientIntensity = scene->ambientRadiance;
  L.sc.occlusionDistance = scene->occlusionDistance > 0.f ? scene->occlusionDistance : 1e20f;
  L.sc.aoSamples = std::min(std::max(scene->ambientSamples, 0), 256);
  L.sc.cullTriangleBF = scene->cullTriangleBackfaces != 0;
  return skip ? launchSceneT<true>(L, s) : launchSceneT<false>(L, s);
}

} // namespace dvr
