/*
 * dvr_b200.h — C-ABI launch layer of the B200-native direct-volume-rendering path.
 *
 * This is the drop-in boundary below the ANARI device (include/anari/anari.h): plain
 * pointers, sizes and POD structs, no C++/torch types.  Every entry point cites the
 * reference (NVIDIA/VisRTX v0.13.0, paths relative to the reference checkout) interface
 * it replaces.  The reference has no C-ABI of its own on this path (it goes host object
 * -> FrameGPUData -> optixLaunch); what a maintainer would bind is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, a negative DvrStatus otherwise; the message of
 *    the last failure on the calling thread is dvr_last_error().
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All
 *    launches are stream-ordered and asynchronous unless stated otherwise.
 *  - there is NO CPU fallback: without a CUDA device every compute entry fails with
 *    DVR_ERR_NO_DEVICE.
 */
#ifndef DVR_B200_H
#define DVR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVR_B200_VERSION_MAJOR 0
#define DVR_B200_VERSION_MINOR 1

#define DVR_TF_SIZE 256        /* TransferFunction1D.h:66 (m_tfDim) */
#define DVR_MACROCELL_WIDTH 16 /* space_skipping/UniformGrid.cu:152-154 */

typedef enum DvrStatus
{
  DVR_OK = 0,
  DVR_ERR_INVALID_ARGUMENT = -1,
  DVR_ERR_NO_DEVICE = -2,
  DVR_ERR_CUDA = -3,
  DVR_ERR_UNSUPPORTED = -4,
  DVR_ERR_OUT_OF_MEMORY = -5
} DvrStatus;

/* voxel element types accepted by a structuredRegular field
 * (StructuredRegularField.cpp:45-60; DVR_F16 is an extension for BASELINE config 3) */
typedef enum DvrDataType
{
  DVR_FLOAT32 = 0,
  DVR_UFIXED8 = 1,
  DVR_FIXED8 = 2,
  DVR_UFIXED16 = 3,
  DVR_FIXED16 = 4,
  DVR_FLOAT64 = 5, /* converted to f32 at upload */
  DVR_FLOAT16 = 6
} DvrDataType;

typedef enum DvrFilter
{
  DVR_FILTER_LINEAR = 0, /* "linear" (default), StructuredRegularField.cpp:93,146 */
  DVR_FILTER_NEAREST = 1
} DvrFilter;

/* colour channel formats: Frame.cu:104-110, gpu_objects.h:625-631 (FrameFormat) */
typedef enum DvrFrameFormat
{
  DVR_FORMAT_FLOAT32_VEC4 = 0,     /* FrameFormat::FLOAT */
  DVR_FORMAT_UFIXED8_VEC4 = 1,     /* FrameFormat::UINT  */
  DVR_FORMAT_UFIXED8_RGBA_SRGB = 2 /* FrameFormat::SRGB  */
} DvrFrameFormat;

typedef enum DvrCameraType
{
  DVR_CAMERA_PERSPECTIVE = 0,
  DVR_CAMERA_ORTHOGRAPHIC = 1
} DvrCameraType;

/* Which integrator the launch runs.  RAYCAST/DEFAULT share the fixed-step marcher of
 * gpu/volumeIntegration.h:64-103; they differ in pixel sampling (Raycast_ptx.cu:67 uses the
 * un-jittered pixel corner, DirectLight_ptx.cu:303 jitters and loops numIterations). */
typedef enum DvrIntegrator
{
  DVR_INTEGRATOR_RAYCAST = 0, /* centred pixel, 1 sample */
  DVR_INTEGRATOR_DEFAULT = 1, /* jittered pixel, spp loop */
  /* `dpt` / `diffuse_pathtracer`: delta (Woodcock) tracking through the majorant grid, isotropic
   * scattering, Russian roulette, ambient light (renderer/DiffusePathTracer_ptx.cu:82-215,
   * gpu/volumeIntegration.h:167-296,352-389, gpu/dda.h:43-121) */
  DVR_INTEGRATOR_DPT = 2,
  /* `test`: no scene access — colour = primary ray direction, depth 1 (renderer/Test_ptx.cu:52-69) */
  DVR_INTEGRATOR_TEST = 3
} DvrIntegrator;

/* ---- opaque device objects ------------------------------------------------------ */
typedef struct DvrField DvrField;   /* replaces StructuredRegularField / NvdbRegularField GPU state */
typedef struct DvrVolume DvrVolume; /* replaces TransferFunction1D GPU state (TF table + majorants) */
typedef struct DvrImage DvrImage;   /* replaces the renderer's background texture (Renderer.cpp:172-198) */

/* ---- POD parameter blocks ------------------------------------------------------- */

/* CameraGPUData, gpu/gpu_objects.h:75-103 */
typedef struct DvrCamera
{
  int32_t type;     /* DvrCameraType */
  float region[4];  /* imageRegion (x0,y0,x1,y1), Camera.cpp:70-72 */
  float pos[3];
  float dir[3];     /* normalised */
  float up[3];      /* normalised */
  /* perspective: dir_du, dir_dv, dir_00; orthographic: pos_du, pos_dv, pos_00 */
  float du[3];
  float dv[3];
  float p00[3];
  float scaledAperture;
  float aspect;
} DvrCamera;

/* one entry of the flattened world: World.cpp:328-335,444-460 + Group volume lists */
typedef struct DvrVolumeInstance
{
  const DvrVolume *volume;
  float worldToObject[12]; /* row-major 3x4 (rows = x,y,z of the object-space point) */
  uint32_t instanceId;     /* instance "id", default ~0u */
  uint32_t _pad;
} DvrVolumeInstance;

/* FrameBuffers, gpu/gpu_objects.h:633-644; every pointer is a DEVICE pointer, NULL = channel off */
typedef struct DvrFrameBuffers
{
  float *colorAccumulation; /* vec4[W*H], required */
  void *outColor;           /* uint32[W*H] or vec4[W*H] by format, required */
  float *depth;
  uint32_t *primId;
  uint32_t *objId;
  uint32_t *instId;
  float *albedo; /* vec3[W*H] accumulation */
  float *normal; /* vec3[W*H] accumulation */
  /* Optional second destination of the encoded colour, same layout as outColor: every pixel the launch writes to
   * outColor is also stored here.  Meant for device-accessible pinned HOST memory (cudaHostAlloc): the frame then
   * reaches the host through posted PCIe writes overlapped with the march instead of a copy after it — what
   * Frame::map("channel.color") needs (frame/Frame.cu:312-330 maps the colour buffer to the host). NULL = off. */
  void *outColorMirror;
  /* Optional second destination of the depth channel (write-only, same layout as depth): every depth value the launch
   * stores is also stored here.  Sort-last: each GPU keeps the depth of the pixels it resolves in its own `depth`
   * (read back by the accumulate step of later frames) and mirrors it into the display GPU's frame through a peer
   * pointer, so no GPU ever LOADS depth over NVLink.  NULL = off. */
  float *depthMirror;
} DvrFrameBuffers;

/* Empty-space skipping never changes a pixel (skipped lattice points classify to alpha == 0 exactly), so the
 * mode is purely a speed choice. */
typedef enum DvrSkipMode
{
  DVR_SKIP_OFF = 0,
  DVR_SKIP_ON = 1,
  DVR_SKIP_AUTO = 2 /* on when >= 1/64 of a volume's macrocells are empty under its transfer function */
} DvrSkipMode;

/* FramebufferGPUData + the RendererGPUData members this path reads
 * (gpu/gpu_objects.h:609-655, Renderer.cpp:191-207) */
typedef struct DvrFrameParams
{
  uint32_t width, height;
  int32_t format;         /* DvrFrameFormat */
  int32_t integrator;     /* DvrIntegrator */
  int32_t frameID;        /* samples already accumulated; 0 => buffers are (re)initialised by the launch */
  int32_t checkerboardID; /* -1 = off, else 0..3 (createScreenSample.h:38-46) */
  int32_t numIterations;  /* pixelSamples (>=1) */
  float inverseVolumeSamplingRate; /* 1/volumeSamplingRate */
  float background[4];    /* constant background colour (Renderer.cpp:154) */
  /* sort-first: only pixels with y in [rowBegin,rowEnd) and tiles owned by this rank are rendered */
  uint32_t tileRank, tileRanks; /* 0,1 = everything */
  /* options of the new implementation (all parity-neutral) */
  int32_t useMacrocellSkipping; /* DvrSkipMode: skip fully transparent macrocells on the same sample lattice */
  int32_t tileBand;             /* sort-first: consecutive tile rows per band owned by one rank (0/1 = finest) */
  /* DVR_INTEGRATOR_DPT only (DiffusePathTracer.cpp:44-54, Renderer.cpp:159-161) */
  int32_t maxDepth;             /* "maxDepth", clamped to [1,256]; 0 => 5 */
  float ambientRadiance;        /* "ambientRadiance" (dpt default 1) */
  float occlusionDistance;      /* "ambientOcclusionDistance"; 0 => 1e20 */
  int32_t dptReferenceGrid;     /* DPT only: != 0 walks the grid built the reference's way (see dvr_volume_dda_majorants) */
  /* dvr_render_partial* only: != 0 renders (and WRITES) only the tiles inside the screen-space rectangle of the
   * volume's bounds; pixels outside keep whatever the partial buffers held.  For consumers that cull by themselves
   * — dvr_composite_resolve_peers* regenerates each primary ray and never reads a pixel whose ray misses. */
  int32_t partialCullToBounds;
  int32_t _reserved[1];
  /* "background" given as an image (Renderer.cpp:154,175-198): non-NULL => every pixel-sample takes its background
   * from a bilinear fetch of the image at its screen coordinate (gpu/gpu_util.h:289-307) and `background` is
   * ignored.  NULL = the constant colour. */
  const DvrImage *backgroundImage;
} DvrFrameParams;

/* per-launch counters, filled only by dvr_render_instrumented (device memory, 64-bit each) */
typedef struct DvrRenderStats
{
  unsigned long long samplesTaken;   /* field fetches (volumeIntegration.h:86-102 loop bodies) */
  unsigned long long samplesSkipped; /* lattice points skipped by macrocell skipping */
  unsigned long long raysHit;        /* pixel-samples that entered at least one volume */
  unsigned long long macrocellsTouched; /* distinct 16^3 cells fetched from (algorithmic bytes) */
} DvrRenderStats;

/* ---- library ------------------------------------------------------------------- */
const char *dvr_last_error(void);
int dvr_version(int *major, int *minor);
/* number of CUDA devices visible; <=0 means the compute entries will fail (VisRTXDevice.cpp:554-578) */
int dvr_device_count(void);
/* cudaSetDevice for the calling thread (the "cudaDevice" device parameter, VisRTXDevice.cpp:464) */
int dvr_set_device(int cudaDevice);
/* name, SM count and memory of the current device */
int dvr_device_info(char *name, size_t nameLen, int *smCount, size_t *totalMem);

/* ---- host-side parameter helpers (pure functions, callable without a GPU) ---------- */

/* Perspective::commitParameters, camera/Perspective.cpp:42-72 + Camera::readBaseParameters :68-76.
 * region may be NULL (=> 0,0,1,1). */
int dvr_camera_perspective(const float pos[3], const float dir[3], const float up[3],
    float fovy, float aspect, float focusDistance, float apertureRadius,
    const float region[4], DvrCamera *out);
/* Orthographic::commitParameters, camera/Orthographic.cpp:38-52 */
int dvr_camera_orthographic(const float pos[3], const float dir[3], const float up[3],
    float height, float aspect, const float region[4], DvrCamera *out);

/* TransferFunction1D::discritizeTFData, scene/volume/TransferFunction1D.cpp:101-150 with
 * generateLinearPositions/getInterpolatedValue, utility/colorMapHelpers.h:43-72.
 * color: nColor entries of colorChannels (3|4) floats, or NULL => uniformColor;
 * opacity: nOpacity floats or NULL => uniformOpacity (already multiplied by uniformColor.a
 * as in TransferFunction1D.cpp:55).  Writes DVR_TF_SIZE rgba texels. */
int dvr_tf_discretize(const float *color, size_t nColor, int colorChannels,
    const float *opacity, size_t nOpacity, const float uniformColor[4], float uniformOpacity,
    const float valueRange[2], float *outRgba);

/* ---- fields ----------------------------------------------------------------------- */

/* StructuredRegularField::finalize, spatial_field/StructuredRegularField.cpp:98-159.
 * data: dims[0]*dims[1]*dims[2] elements, x fastest; host pointer or device pointer
 * (dataIsDevice != 0; ANARI_NV_ARRAY_CUDA, array/Array.cpp:38-66).  The voxels are copied into a
 * 3-D CUDA array bound to a clamp/normalised texture exactly as the reference does; the
 * caller's buffer is not referenced afterwards.  Synchronous w.r.t. the host when data is
 * pageable host memory. */
int dvr_field_create_structured(const void *data, int dataIsDevice, int dataType /*DvrDataType*/,
    const uint32_t dims[3], const float origin[3], const float spacing[3],
    int filter /*DvrFilter*/, void *stream, DvrField **out);
/* Same, but only z-slices [zBegin, zEnd) (+ ghost layers clamped to the volume) of a global
 * dims[] volume are resident: the sort-last slab of SURVEY 8e.  data points at the first
 * RESIDENT slice (ghost included): slice index max(zBegin-1,0).  data == NULL allocates only (see
 * dvr_field_upload_slices). */
int dvr_field_create_structured_slab(const void *data, int dataIsDevice, int dataType,
    const uint32_t globalDims[3], uint32_t zBegin, uint32_t zEnd, const float origin[3],
    const float spacing[3], int filter, void *stream, DvrField **out);
/* Moves the ownership of a slab field without moving a voxel: afterwards the field takes the lattice samples whose cell
 * slice lies in [zBegin, zEnd).  The new range must lie inside the range the slab was created with (its resident slices
 * are those +- one ghost slice), so a slab created with a margin of M slices on each side can hand up to M slices to
 * either neighbour — and take them back — between two frames.  Sort-last load balancing: the per-GPU march times of the
 * last frames decide where the cuts go (visrtx_b200/multigpu.py rebalance_slab_ranges); every GPU of the partition must
 * apply the same cuts before the next frame.  Host-side bookkeeping only (the kernels read the range per launch). */
int dvr_field_set_owned_slices(DvrField *f, uint32_t zBegin, uint32_t zEnd);
/* current ownership and the limits it can move in (the range given at creation) */
int dvr_field_owned_slices(const DvrField *f, uint32_t *zBegin, uint32_t *zEnd, uint32_t *limitBegin, uint32_t *limitEnd);
/* Chunked fill of a slab field created with data == NULL (volumes too large to stage twice in HBM):
 * copies nSlices z-slices into resident slices [firstResidentSlice, +nSlices); element type = the
 * field's.  Call dvr_field_build_macrocells once all slices are in. */
int dvr_field_upload_slices(DvrField *f, const void *data, int dataIsDevice, uint32_t firstResidentSlice,
    uint32_t nSlices, void *stream);
/* Re-finalisation of a whole structuredRegular field whose `data` changed while dims, element type and filter
 * did not (time-varying / in-situ fields).  The reference re-runs StructuredRegularField::finalize on every
 * commit of the field — cleanup, cudaMalloc3DArray, copy, grid rebuild (StructuredRegularField.cpp:98-159) —
 * here the 3-D array, its texture views and the macrocell storage are kept: one copy into the array and one
 * macrocell build (straight from `data` when that is linear f32 device memory).  origin / spacing may change.
 * The caller then calls dvr_volume_update on every volume bound to the field (its majorants derive from the
 * ranges).  DVR_ERR_UNSUPPORTED for slabs, NanoVDB, FLOAT64 or a different dataType: destroy + create instead. */
int dvr_field_update_structured(DvrField *f, const void *data, int dataIsDevice, int dataType,
    const float origin[3], const float spacing[3], void *stream);
/* NvdbRegularField::finalize, spatial_field/NvdbRegularField.cpp:64-160: `gridData` is one serialized
 * NanoVDB grid (the ANARI "nanovdb" field's UINT8 `data` array); GridType Float, Fp4, Fp8, Fp16 and FpN
 * (sampleSpatialField.h:107-109), anything else -> DVR_ERR_UNSUPPORTED.  The
 * buffer is copied to 32-byte aligned device memory; bounds = the grid's world bounding box, step =
 * min(voxelSize)/2; macrocells cover the index bounding box. */
int dvr_field_create_nanovdb(const void *gridData, size_t bytes, int dataIsDevice, void *stream, DvrField **out);
int dvr_field_destroy(DvrField *f);
/* SpatialField::bounds / stepSize, StructuredRegularField.cpp:166-178 */
int dvr_field_bounds(const DvrField *f, float lower[3], float upper[3]);
int dvr_field_step_size(const DvrField *f, float *stepSize);
/* bytes of device memory held by the field (voxels + macrocells) */
int dvr_field_device_bytes(const DvrField *f, size_t *bytes);

/* UniformGrid::init + buildGrid, space_skipping/UniformGrid.cu:152-224, rebuilt as ONE pass over
 * the voxels that records the conservative [min,max] of every value a trilinear fetch inside
 * the 16^3 cell can return (cell voxels + one-voxel apron); see DESIGN.md (reference quirk Q7
 * is deliberately not reproduced).  Called implicitly by dvr_field_create_*. */
int dvr_field_build_macrocells(DvrField *f, void *stream);
/* grid dims (ceil(dims/16)) and device pointers: valueRanges = float2[n] (lower,upper) */
int dvr_field_macrocells(const DvrField *f, uint32_t gridDims[3], const float **valueRangesDev);
/* global scalar range of the field (min,max), reduced from the macrocell ranges
 * (tsd/src/tsd/algorithms/computeScalarRange.cpp).  Synchronises the stream. */
int dvr_field_value_range(const DvrField *f, void *stream, float range[2]);

/* ---- background image ------------------------------------------------------------- */

/* component types of the image array the application hands to the renderer's "background" parameter */
typedef enum DvrImageComponent
{
  DVR_IMAGE_FLOAT32 = 0, /* ANARI_FLOAT32[_VECn] */
  DVR_IMAGE_UFIXED8 = 1, /* ANARI_UFIXED8[_VECn] */
  DVR_IMAGE_UFIXED16 = 2,
  DVR_IMAGE_UFIXED32 = 3,
  DVR_IMAGE_SRGB8 = 4    /* ANARI_UFIXED8_R[GBA]_SRGB: linearised at upload */
} DvrImageComponent;

/* Renderer::finalize (renderer/Renderer.cpp:172-179): Array2D::acquireCUDAArrayUint8 -> makeCudaArrayUint8
 * (utility/CudaImageTexture.cpp:43-58,84-101,139-226) + makeCudaTextureObject(array, normalizedFloat, "linear")
 * (:316-345).  `pixels` is HOST memory, width*height elements of `channels` (1..4) components, row-major from the
 * bottom row (row 0 is sampled at screen y = 0).  Every component is converted to 8 bits exactly as the reference's
 * staging pass does (float: uint8(c*255), truncating; 16/32-bit fixed: uint8(c/max*255); sRGB8: linearised), three
 * channels are padded with alpha 255, and the result is bound to a clamp / linear / normalised-coordinate texture. */
int dvr_image_create(const void *pixels, int componentType /*DvrImageComponent*/, int channels, uint32_t width,
    uint32_t height, void *stream, DvrImage **out);
int dvr_image_destroy(DvrImage *img);

/* ---- volumes ---------------------------------------------------------------------- */

/* TransferFunction1D::finalize/gpuData, TransferFunction1D.cpp:66-99,152-186: uploads the 256-texel
 * table, stores valueRange / 1/unitDistance / id, and computes per-macrocell majorants
 * (UniformGrid::computeMaxOpacities, UniformGrid.cu:55-90,248-258 — with the volume's own
 * valueRange, i.e. quirk Q8 fixed).  tfRgba = DVR_TF_SIZE*4 floats on the HOST. */
int dvr_volume_create(const DvrField *field, const float *tfRgba, const float valueRange[2],
    float unitDistance, uint32_t id, void *stream, DvrVolume **out);
/* re-upload after a transfer-function edit (same semantics as a re-finalize) */
int dvr_volume_update(DvrVolume *v, const float *tfRgba, const float valueRange[2],
    float unitDistance, uint32_t id, void *stream);
int dvr_volume_destroy(DvrVolume *v);
/* device pointer to float[nMacrocells] majorants (max TF alpha over the cell's value range) */
int dvr_volume_majorants(const DvrVolume *v, const float **maxOpacitiesDev);
/* The delta-tracking grid the dpt integrator walks (UniformGridData of gpu/gpu_objects.h:395-401 as
 * filled by UniformGrid::init/buildGrid/computeMaxOpacities, UniformGrid.cu:143-258): dims = ceil(field
 * dims / 16) cells dividing the field bounds evenly; *maxOpacitiesDev = device float[dims.x*dims.y*dims.z].
 * Built on first use (this call or the first DVR_INTEGRATOR_DPT frame) and after dvr_volume_update.
 * referenceBuild == 0 (default): conservative content (value ranges over every voxel a trilinear stencil in the
 * cell can touch, classified with the volume's own value range).  referenceBuild != 0: the content the
 * reference itself computes — buildGridGPU's one sample octet per macrocell at [0,1] coordinates and
 * computeMaxOpacities with the default {0,1} range — for bit-level comparison with VisRTX. */
int dvr_volume_dda_majorants(DvrVolume *v, int32_t referenceBuild, void *stream, uint32_t dims[3],
    const float **maxOpacitiesDev);

/* ---- the hot path ------------------------------------------------------------------- */

/* One frame launch: replaces Frame::renderFrame's newFrame()+upload()+optixLaunch
 * (frame/Frame.cu:272-289) and everything the raygen program does per pixel
 * (renderer/Raycast_ptx.cu:60-179, DirectLight_ptx.cu:294-418 volume branch,
 * gpu/volumeIntegration.h:64-165,317-350, scene/Intersectors_ptx.cu:248-274,
 * gpu/gpu_util.h:393-443).  frameID==0 (re)initialises accumulation/depth/id buffers inside
 * the launch, replacing the thrust::fill_n calls of Frame.cu:609-647. */
int dvr_render(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instances, uint32_t nInstances, const DvrFrameBuffers *buffers,
    void *stream);
/* Same launch with counters (slower; for bench bookkeeping and tests, never timed).
 * statsDev: device pointer to one DvrRenderStats, zeroed by the call. */
int dvr_render_instrumented(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instances, uint32_t nInstances, const DvrFrameBuffers *buffers,
    DvrRenderStats *statsDev, void *stream);
/* number of kernel launches issued by this library since load (bench.py "gpu_launches") */
unsigned long long dvr_launch_count(void);

/* ---- surfaces in front of / behind volumes, lights and shadow rays ---------------------------
 * The mixed-scene branch of the `default` (directLight) and `raycast` renderers: primary rays find the closest
 * surface, volumes are marched up to surfaceHit.t, the surface is shaded (matte, ambient + direct light with
 * surface- and volume-attenuated shadow rays, ambient occlusion) and the loop continues behind a translucent
 * surface (renderer/DirectLight_ptx.cu:64-218,294-418, Raycast_ptx.cu:60-179, gpu/computeAO.h:39-60,
 * gpu/sampleLight.h:54-80, shaders/MatteShader_ptx.cu:39-80).  OptiX's RT-core traversal is replaced by a BVH over
 * all primitives of the flattened world, built on the host at dvr_surfaces_create. */
typedef struct DvrSurfaces DvrSurfaces; /* replaces the surface TLAS/BLAS + SurfaceGPUData / GeometryGPUData / MaterialGPUData */

typedef enum DvrGeometryType
{
  DVR_GEOMETRY_TRIANGLE = 0, /* scene/surface/geometry/Triangle.cpp: vertex.position, primitive.index, vertex.normal */
  DVR_GEOMETRY_SPHERE = 1    /* scene/surface/geometry/Sphere.cu: vertex.position, primitive.index, vertex.radius, radius */
} DvrGeometryType;

typedef enum DvrAlphaMode /* gpu/evalMaterialParameters.h:392-403 */
{
  DVR_ALPHA_OPAQUE = 0,
  DVR_ALPHA_BLEND = 1,
  DVR_ALPHA_MASK = 2
} DvrAlphaMode;

/* One surface instance of the flattened world: geometry + matte material + instance transform.  All arrays are HOST
 * memory and are copied by dvr_surfaces_create. */
typedef struct DvrSurfaceDesc
{
  int32_t geometryType;        /* DvrGeometryType */
  uint32_t nVertices;
  const float *vertexPosition; /* vec3[nVertices]: triangle corners / sphere centres */
  uint32_t nPrimitives;        /* triangles / spheres; with index == NULL: nVertices / 3 resp. nVertices */
  const uint32_t *index;       /* triangle: uvec3[nPrimitives]; sphere: uint32[nPrimitives]; NULL = soup */
  const float *vertexNormal;   /* triangle: vec3[nVertices] or NULL (=> geometric normal) */
  const float *vertexRadius;   /* sphere: float[nVertices] or NULL (=> radius) */
  float radius;                /* sphere "radius" (default 0.01) */
  const uint32_t *primitiveId; /* "primitive.id": uint32[nPrimitives] or NULL (=> primitive index) */
  int32_t cullBackfaces;       /* triangle "cullBackfaces" (gpu/populateHit.h:331-343) */
  /* matte material (scene/surface/material/Matte.cpp:38-52): constant colour / opacity only */
  float color[4];              /* default (0.8, 0.8, 0.8, 1) */
  float opacity;               /* default 1 */
  int32_t alphaMode;           /* DvrAlphaMode, default opaque */
  float alphaCutoff;           /* default 0.5 */
  uint32_t surfaceId;          /* surface "id", default ~0u */
  uint32_t instanceId;         /* instance "id", default ~0u */
  float objectToWorld[12];     /* row-major 3x4 */
} DvrSurfaceDesc;

int dvr_surfaces_create(const DvrSurfaceDesc *surfaces, uint32_t nSurfaces, void *stream, DvrSurfaces **out);
int dvr_surfaces_destroy(DvrSurfaces *s);
/* primitives and BVH nodes held (tests / bookkeeping) */
int dvr_surfaces_info(const DvrSurfaces *s, uint32_t *nPrimitives, uint32_t *nNodes);

typedef enum DvrLightType
{
  DVR_LIGHT_DIRECTIONAL = 0, /* scene/light/Directional.cpp: direction (normalised), irradiance */
  DVR_LIGHT_POINT = 1        /* scene/light/Point.cpp: position, intensity (or power) */
} DvrLightType;

/* one light instance, already transformed by its instance (gpu/sampleLight.h:54-78 applies xfmVec / xfmPoint) */
typedef struct DvrLight
{
  int32_t type;
  float color[3];
  float vec[3];   /* directional: world-space direction the light travels in; point: world-space position */
  float strength; /* irradiance / intensity */
} DvrLight;

/* the RendererGPUData members the surface branch reads (Renderer.cpp:152-207, DirectLight.cpp:49-64) */
typedef struct DvrSceneParams
{
  const DvrSurfaces *surfaces; /* NULL or empty: dvr_render_scene == dvr_render */
  const DvrLight *lights;      /* HOST array */
  uint32_t nLights;
  float ambientColor[3];       /* "ambientColor", default 1,1,1 */
  float ambientRadiance;       /* "ambientRadiance" (directLight default 0) */
  float occlusionDistance;     /* "ambientOcclusionDistance"; 0 => 1e20 */
  int32_t ambientSamples;      /* "ambientSamples" clamped to [0,256], default 1 */
  int32_t cullTriangleBackfaces; /* "cullTriangleBackfaces" */
} DvrSceneParams;

/* dvr_render for a world that also holds surfaces and lights.  DVR_INTEGRATOR_DEFAULT runs the directLight raygen
 * (jittered pixel, numIterations), DVR_INTEGRATOR_RAYCAST the raycast one (|dir . Ns| * ambientColor headlight); any
 * other integrator -> DVR_ERR_UNSUPPORTED.  Volume segments use the same marcher as dvr_render, so a pixel whose
 * rays meet no surface is bit-identical to dvr_render. */
int dvr_render_scene(const DvrFrameParams *params, const DvrCamera *camera, const DvrVolumeInstance *instances,
    uint32_t nInstances, const DvrSceneParams *scene, const DvrFrameBuffers *buffers, void *stream);

/* ---- sort-last (slab) rendering and compositing, SURVEY 8e ---------------------------- */

/* Partial render of the slab fields on the GLOBAL sample lattice: writes premultiplied
 * (C, A) as float4 and the entry depth per pixel-sample, WITHOUT `color *= opacity`,
 * background, tonemap or accumulation (those run once in dvr_resolve on the composited value).
 * partialRgba: float4[W*H]; partialDepth: float[W*H]. Only DVR_INTEGRATOR_RAYCAST|DEFAULT with
 * numIterations==1 and a single volume instance. */
int dvr_render_partial(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instance, float *partialRgba, float *partialDepth, void *stream);
/* same launch with counters (samplesTaken/Skipped, macrocellsTouched of this slab); never timed */
int dvr_render_partial_instrumented(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instance, float *partialRgba, float *partialDepth, DvrRenderStats *statsDev,
    void *stream);
/* front-to-back `over` of two partial images for pixels [pixelBegin,pixelEnd):
 * front = front over back (in place on front); depth = min.  Pointers may be peer-mapped. */
int dvr_composite_over(float *frontRgba, float *frontDepth, const float *backRgba,
    const float *backDepth, size_t pixelBegin, size_t pixelEnd, int backIsInFront, void *stream);
/* Resolve a composited partial image: applies Raycast_ptx.cu:159-165 (color*=opacity, background)
 * and gpu_util.h:393-443 (accumulate, tonemap, encode) for pixels [pixelBegin,pixelEnd). */
int dvr_resolve(const DvrFrameParams *params, const float *partialRgba, const float *partialDepth,
    uint32_t objId, uint32_t instId, const DvrFrameBuffers *buffers, size_t pixelBegin,
    size_t pixelEnd, void *stream);

/* Direct-send compositing fused with the resolve, over peer memory: for pixels [pixelBegin,pixelEnd)
 * the nSlabs partial images (device pointers, possibly CUDA-IPC / peer mapped, given in ascending z
 * order of their slabs) are combined front-to-back with `over` — the per-pixel view order is the sign
 * of the primary ray's z direction — and the result goes through the same tail as dvr_resolve.
 * buffers->outColor may itself be a peer pointer (the display GPU's frame).  One launch per rank
 * replaces log2(N) binary-swap rounds + gather. */
int dvr_composite_resolve_peers(const DvrFrameParams *params, const DvrCamera *camera,
    const float *const *partialRgba, const float *const *partialDepth, uint32_t nSlabs, uint32_t objId,
    uint32_t instId, const DvrFrameBuffers *buffers, size_t pixelBegin, size_t pixelEnd, void *stream);

/* Device-side cross-GPU ordering without host or NCCL involvement.  A flag is a 32-bit word in
 * (CUDA-IPC shared) device memory holding the number of the last frame its producer completed.
 *   signal[i]  : words written with `value` when the launch has finished all its work (system-scope
 *                release by the last retiring warp / block) — typically one word in every peer's flag table
 *   wait       : local flag table; the launch does not touch peer data before wait[i] >= waitValue for all
 *                i < nWait (bounded spin, ~2 s, then the launch gives up and sets *errorFlag if non-NULL)
 * Used by the *_sync variants below and dvr_wait_flags. */
typedef struct DvrPeerSync
{
  uint32_t nSignal;
  uint32_t signalValue;
  unsigned int *signal[16];
  uint32_t nWait;
  uint32_t waitValue;
  const unsigned int *wait;
  unsigned int *errorFlag;
} DvrPeerSync;

/* dvr_render_partial + signal when the partial image is complete */
int dvr_render_partial_sync(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instance, float *partialRgba, float *partialDepth, const DvrPeerSync *sync,
    void *stream);
/* dvr_composite_resolve_peers that (a) waits for every slab's "partial complete" flag inside the kernel,
 * (b) skips the peer loads of pixels whose primary ray (regenerated with the same Philox stream as the
 * partial march) misses the volume's bounds, (c) signals "strip resolved" at the end.  instance = the
 * volume instance whose slabs are being composited (bounds + transform for (b)). */
int dvr_composite_resolve_peers_sync(const DvrFrameParams *params, const DvrCamera *camera,
    const DvrVolumeInstance *instance, const float *const *partialRgba, const float *const *partialDepth,
    uint32_t nSlabs, uint32_t objId, uint32_t instId, const DvrFrameBuffers *buffers, size_t pixelBegin,
    size_t pixelEnd, const DvrPeerSync *sync, void *stream);
/* one-thread kernel: returns (in stream order) once flags[i] >= value for all i < n (bounded spin) */
int dvr_wait_flags(const unsigned int *flags, uint32_t n, uint32_t value, unsigned int *errorFlag, void *stream);

/* The whole sort-last frame of one GPU in ONE launch: march of the own slab, exchange and compositing fused.
 * Tiles of the volume's screen window are marched in the same order on every GPU and grouped into regions; a GPU that
 * completes a region publishes a flag to the region's owner (region mod nRanks), and warps that run out of march tiles
 * composite the regions their GPU owns as soon as all nRanks flags of a region are there — peer loads of the partial
 * pixels over NVLink, `over` in per-pixel view order, Q2 + background + accumulate + tonemap + encode, stores into
 * `buffers` (outColor / depth / ids may be peer pointers into the display GPU's frame; colorAccumulation is local:
 * every pixel is always resolved by the same GPU while the camera stands still).  The pixels outside the window are
 * shared out in contiguous strips.  The exchange overlaps the march region by region instead of following it
 * (replaces dvr_render_partial_sync + dvr_composite_resolve_peers_sync + their wait / signal launches).
 *
 * All tables live in (CUDA-IPC or peer-mapped) device memory and are zero before the first frame:
 *   regionFlags[p]   rank p's table uint32[maxRegions][16]: entry [r][q] = last frame in which rank q finished region r
 *   resolvedFlags[p] rank p's table uint32[16]: entry [q] = last frame rank q finished compositing
 *   regionDone       LOCAL scratch uint32[maxRegions] (tile counters)
 * partialRgba[q] / partialDepth[q]: rank q's partial image of THIS frame (float4[W*H] / float[W*H], ascending z order of
 * the slabs; entry [rank] is the one this launch writes).  Alternate two sets of partial images between frames.
 * seq: frame number, strictly increasing from 1, identical on all ranks.  waitAllResolved != 0 (display GPU): the
 * launch completes only when every rank's pixels of this frame have landed. */
typedef struct DvrSlabExchange
{
  uint32_t nRanks, rank;
  uint32_t seq;
  uint32_t maxRegions;
  const float *const *partialRgba;
  const float *const *partialDepth;
  unsigned int *const *regionFlags;
  unsigned int *const *resolvedFlags;
  unsigned int *regionDone;
  unsigned int *errorFlag; /* set to 1 when a bounded spin (~2 s) gave up; may be NULL */
  int32_t waitAllResolved;
  int32_t _pad;
  /* optional (bench bookkeeping): device uint64[8], %globaltimer nanoseconds of THIS GPU — [0] first CTA started
   * (caller presets ~0), [1] last march tile finished, [2] last background chunk finished, [3] last owned region
   * composited, [4] retire (display GPU: after every rank's flag arrived), [5] last owned region seen complete, [6] last
   * region flag published; [1..7] preset 0.  NULL = off. */
  unsigned long long *timing;
} DvrSlabExchange;

int dvr_render_slab_frame(const DvrFrameParams *params, const DvrCamera *camera, const DvrVolumeInstance *instance,
    uint32_t objId, uint32_t instId, const DvrFrameBuffers *buffers, const DvrSlabExchange *exchange, void *stream);

/* ---- CUDA IPC plumbing for one-process-per-GPU sharing of frame / partial buffers ------------- */
#define DVR_IPC_HANDLE_BYTES 64
int dvr_ipc_alloc(size_t bytes, void **devPtr, unsigned char handle[DVR_IPC_HANDLE_BYTES]);
int dvr_ipc_open(const unsigned char handle[DVR_IPC_HANDLE_BYTES], void **devPtr);
int dvr_ipc_close(void *devPtr);
int dvr_ipc_free(void *devPtr);

/* ---- map-time helpers ------------------------------------------------------------------ */
/* Frame::mapAlbedoBuffer / mapNormalBuffer, frame/Frame.cu:521-557: out = accum * invFrameID */
int dvr_scale_vec3(const float *accumVec3, float *outVec3, size_t nPixels, float scale, void *stream);

/* ---- host helper ------------------------------------------------------------------------------------
 * The conservative pixel rectangle [x0,x1) x [y0,y1) outside which no primary ray of `camera` (jitter included) can
 * hit the axis-aligned box: the projection of the 8 corners through the camera model of gpu/cameraCreateRay.h:38-81
 * (perspective / orthographic, image region), padded by 2 pixels.  The frame kernel gives pixels outside it the
 * background without ray set-up, and the sort-last partial march skips them.  Pure host arithmetic (no GPU).
 * Returns 1 and the rectangle, or 0 and the whole frame when no bound can be trusted (thin-lens camera, a corner
 * beside or behind the eye, degenerate camera basis); < 0 on bad arguments. */
int dvr_bounds_screen_rect(const DvrCamera *camera, const float boundsLo[3], const float boundsHi[3], uint32_t width,
    uint32_t height, int32_t rect[4]);

/* ---- self-test ------------------------------------------------------------------------------------
 * Empty-space skipping advances a ray over n lattice points in closed form (per float binade) instead of n
 * dependent `t += step` additions; the lattice must stay bit-identical to the reference's loop
 * (gpu/volumeIntegration.h:86-102).  Runs both on `count` pseudo-random operand sets on the device and returns
 * the number that differ (must be 0). */
int dvr_selftest_lattice_advance(uint32_t count, uint64_t seed, uint32_t *mismatchesOut, void *stream);

/* ---- frame post passes on device buffers ----------------------------------
 * What TSD's render pipeline runs on the channels it maps through ANARI_NV_FRAME_BUFFERS_CUDA
 * (the .cpp files of tsd/src/render_pipeline/passes).  All pointers are device pointers (the `...CUDA` channel maps or
 * pipeline buffers); launches are stream-ordered. */
/* convertFloatColorBuffer, AnariSceneRenderPass.cpp:15-20: uint8(clamp(v,0,1) * 255) per component */
int dvr_post_convert_float_color(const float *rgbaF32, uint32_t *rgba8, size_t nPixels, void *stream);
/* compositeFrame, AnariSceneRenderPass.cpp:30-46: take the incoming pixel when firstPass or when it is closer;
 * idIn / idOut may be NULL together */
int dvr_post_composite_depth(uint32_t *colorOut, float *depthOut, uint32_t *idOut, const uint32_t *colorIn,
    const float *depthIn, const uint32_t *idIn, size_t nPixels, int firstPass, void *stream);
/* computeOutline + shadePixel, OutlineRenderPass.cpp:13-46: a pixel whose 3x3 neighbourhood holds 2..7 pixels of
 * objectId == outlineId is blended 80 % towards orange (1, .5, 0).  The reference's unsigned `max(0u, y - 1)`
 * wraps on row / column 0, so those never get an outline — reproduced. */
int dvr_post_outline(uint32_t *rgba8, const uint32_t *objectId, uint32_t width, uint32_t height, uint32_t outlineId,
    void *stream);
/* computeDepthImage, VisualizeDepthPass.cpp:13-21: grey = clamp(depth / maxDepth, 0, 1), alpha 1 */
int dvr_post_visualize_depth(uint32_t *rgba8, const float *depth, size_t nPixels, float maxDepth, void *stream);
/* pick operation of PickPass.cpp: what is under pixel (x, y) — copies one depth and one id to the host
 * (synchronises the stream); objectId may be NULL (id = ~0u) */
int dvr_post_pick(const float *depth, const uint32_t *objectId, uint32_t width, uint32_t height, uint32_t x, uint32_t y,
    float *depthOut, uint32_t *idOut, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DVR_B200_H */
