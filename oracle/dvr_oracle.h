/* dvr_oracle.h — C interface of O-cpu, the CPU restatement of the reference DVR path.
 * TEST INFRASTRUCTURE: loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Uses the POD structs of include/dvr_b200.h for the
 * camera and frame parameters so both sides are driven by identical inputs. */
#ifndef DVR_ORACLE_H
#define DVR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/dvr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleVolume
{
  const float *voxels; /* dims[0]*dims[1]*dims[2] floats, x fastest (fixed-point types pre-normalised) */
  int32_t dims[3];
  float origin[3];
  float spacing[3];
  int32_t filterNearest;
  const float *tf; /* 256 rgba texels */
  float valueRange[2];
  float unitDistance;
  uint32_t id;
  float worldToObject[12];
  uint32_t instanceId;
  int32_t zOwnBegin, zOwnEnd; /* 0,0 = whole volume; otherwise only these cell slices are sampled */
  const void *nvdbGrid;       /* non-NULL: a serialized NanoVDB float grid ("nanovdb" field); voxels/dims unused */
} OracleVolume;

typedef struct OracleBuffers
{
  float *colorAccumulation;
  void *outColor;
  float *depth;
  uint32_t *primId, *objId, *instId;
  float *albedo, *normal;
} OracleBuffers;

int oracle_camera_perspective(const float pos[3], const float dir[3], const float up[3], float fovy, float aspect,
    float focusDistance, float apertureRadius, const float region[4], DvrCamera *out);
int oracle_camera_orthographic(const float pos[3], const float dir[3], const float up[3], float height,
    float aspect, const float region[4], DvrCamera *out);
int oracle_tf_discretize(const float *color, size_t nColor, int colorChannels, const float *opacity,
    size_t nOpacity, const float uniformColor[4], float uniformOpacity, const float valueRange[2], float *outRgba);
float oracle_nvdb_sample(const void *grid, float x, float y, float z);
float oracle_tex3d(const float *voxels, const int dims[3], float u, float v, float w);
void oracle_tex1d_tf(const float *tf, float coord, float out[4]);
void oracle_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_philox_uniforms(uint64_t seed, uint64_t offset, int n, float *out);
/* renders launch rows [rowBegin,rowEnd) (0,0 = all); samplesOut = field fetches */
int oracle_render(const DvrFrameParams *params, const DvrCamera *camera, const OracleVolume *volumes, int nVolumes,
    const OracleBuffers *buffers, uint64_t *samplesOut, int rowBegin, int rowEnd);

/* background image of the following oracle_render calls (gpu_util.h:289-296); texels = w*h*channels bytes (channels
 * 1, 2 or 4: the RGBA8 texture the renderer builds, Renderer.cpp:172-179), NULL = back to the constant colour */
int oracle_set_background_image(const uint8_t *texels, int channels, int w, int h);

/* dpt renderer: the delta-tracking grid (ceil(dims/16) cells over the bounds) with conservative majorants */
int oracle_dda_majorants(const OracleVolume *volume, int32_t dims[3], float *out, size_t capacity);

#ifdef __cplusplus
}
#endif
#endif
