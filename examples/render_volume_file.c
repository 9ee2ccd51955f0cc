/* render_volume_file.c — a plain-C ANARI application on the drop-in device, in the shape of the reference's
 * examples/simple/testApp_spheres.cpp:167-263 and tsd::import_volume: load a volume file (.raw / .mhd / .vti /
 * .nvdb) with the importers of include/dvr_import.h, build field -> transferFunction1D volume -> world, orbit
 * camera, render N progressive frames and write the colour channel as a binary PPM.
 *
 *   cc -std=c99 -Iinclude examples/render_volume_file.c -Lvisrtx_b200 -lanari_library_visrtx_b200 -ldvr_import \
 *      -Wl,-rpath,$PWD/visrtx_b200 -lm -o render_volume_file
 *   ./render_volume_file volume_64x64x64_uint8.raw out.ppm [renderer=default] [frames=8] [width=512] [height=512]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "anari/anari.h"
#include "anari/ext/visrtx_b200.h"
#include "dvr_b200.h"
#include "dvr_import.h"

static void statusCb(const void *userPtr, ANARIDevice dev, ANARIObject src, ANARIDataType srcType,
    ANARIStatusSeverity sev, ANARIStatusCode code, const char *msg)
{
  (void)userPtr, (void)dev, (void)src, (void)srcType, (void)code;
  if (sev <= ANARI_SEVERITY_WARNING)
    fprintf(stderr, "[anari %d] %s\n", (int)sev, msg);
}

static ANARIDataType elementType(int dvrType)
{
  switch (dvrType) {
  case DVR_UFIXED8: return ANARI_UFIXED8;
  case DVR_FIXED8: return ANARI_FIXED8;
  case DVR_UFIXED16: return ANARI_UFIXED16;
  case DVR_FIXED16: return ANARI_FIXED16;
  case DVR_FLOAT64: return ANARI_FLOAT64;
  default: return ANARI_FLOAT32;
  }
}

int main(int argc, char **argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: %s <volume file> <out.ppm> [renderer] [frames] [width] [height]\n", argv[0]);
    return 2;
  }
  const char *renderer = argc > 3 ? argv[3] : "default";
  const int frames = argc > 4 ? atoi(argv[4]) : 8;
  const uint32_t size[2] = {argc > 5 ? (uint32_t)atoi(argv[5]) : 512u, argc > 6 ? (uint32_t)atoi(argv[6]) : 512u};

  DvrVolumeFile vf;
  if (dvr_import_volume(argv[1], &vf) != DVR_IMPORT_OK) {
    fprintf(stderr, "%s\n", dvr_import_last_error());
    return 1;
  }

  ANARIDevice d = makeVisRTXDevice(statusCb, NULL);
  if (!d)
    return 1;

  /* spatial field */
  ANARISpatialField field;
  if (vf.kind == DVR_IMPORT_STRUCTURED) {
    ANARIArray3D data = anariNewArray3D(d, vf.data, NULL, NULL, elementType(vf.dataType), vf.dims[0], vf.dims[1], vf.dims[2]);
    field = anariNewSpatialField(d, "structuredRegular");
    anariSetParameter(d, field, "data", ANARI_ARRAY3D, &data);
    anariSetParameter(d, field, "origin", ANARI_FLOAT32_VEC3, vf.origin);
    anariSetParameter(d, field, "spacing", ANARI_FLOAT32_VEC3, vf.spacing);
    anariCommitParameters(d, field);
    anariRelease(d, data);
  } else {
    ANARIArray1D data = anariNewArray1D(d, vf.data, NULL, NULL, ANARI_UINT8, vf.bytes);
    field = anariNewSpatialField(d, "nanovdb");
    anariSetParameter(d, field, "data", ANARI_ARRAY1D, &data);
    anariCommitParameters(d, field);
    anariRelease(d, data);
  }

  /* volume: the TSD default colour map (red -> green -> blue, alpha 0 -> .5 -> 1) */
  const float cmap[3][4] = {{1.f, 0.f, 0.f, 0.f}, {0.f, 1.f, 0.f, .5f}, {0.f, 0.f, 1.f, 1.f}};
  ANARIArray1D color = anariNewArray1D(d, cmap, NULL, NULL, ANARI_FLOAT32_VEC4, 3);
  ANARIVolume volume = anariNewVolume(d, "transferFunction1D");
  anariSetParameter(d, volume, "value", ANARI_SPATIAL_FIELD, &field);
  anariSetParameter(d, volume, "color", ANARI_ARRAY1D, &color);
  anariSetParameter(d, volume, "valueRange", ANARI_FLOAT32_BOX1, vf.valueRange);
  float unitDistance = 4.f * fminf(fminf(vf.spacing[0], vf.spacing[1]), vf.spacing[2]);
  anariSetParameter(d, volume, "unitDistance", ANARI_FLOAT32, &unitDistance);
  anariCommitParameters(d, volume);
  anariRelease(d, color);

  ANARIWorld world = anariNewWorld(d);
  ANARIArray1D volumes = anariNewArray1D(d, &volume, NULL, NULL, ANARI_VOLUME, 1);
  anariSetParameter(d, world, "volume", ANARI_ARRAY1D, &volumes);
  anariCommitParameters(d, world);
  anariRelease(d, volumes);

  /* orbit camera around the world bounds (tsd/apps/tools/tsdRender.cpp:183-200) */
  float b[6];
  anariGetProperty(d, world, "bounds", ANARI_FLOAT32_BOX3, b, sizeof(b), ANARI_WAIT);
  const float c[3] = {.5f * (b[0] + b[3]), .5f * (b[1] + b[4]), .5f * (b[2] + b[5])};
  const float diag = sqrtf((b[3] - b[0]) * (b[3] - b[0]) + (b[4] - b[1]) * (b[4] - b[1]) + (b[5] - b[2]) * (b[5] - b[2]));
  const float az = 30.f * 3.14159265f / 180.f, el = 20.f * 3.14159265f / 180.f, dist = 1.2f * diag;
  const float eye[3] = {c[0] + sinf(az) * cosf(el) * dist, c[1] + sinf(el) * dist, c[2] + cosf(az) * cosf(el) * dist};
  float dir[3] = {c[0] - eye[0], c[1] - eye[1], c[2] - eye[2]};
  const float dl = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  dir[0] /= dl, dir[1] /= dl, dir[2] /= dl;
  const float up[3] = {0.f, 1.f, 0.f};
  const float fovy = 60.f * 3.14159265f / 180.f, aspect = (float)size[0] / (float)size[1];
  ANARICamera camera = anariNewCamera(d, "perspective");
  anariSetParameter(d, camera, "position", ANARI_FLOAT32_VEC3, eye);
  anariSetParameter(d, camera, "direction", ANARI_FLOAT32_VEC3, dir);
  anariSetParameter(d, camera, "up", ANARI_FLOAT32_VEC3, up);
  anariSetParameter(d, camera, "fovy", ANARI_FLOAT32, &fovy);
  anariSetParameter(d, camera, "aspect", ANARI_FLOAT32, &aspect);
  anariCommitParameters(d, camera);

  ANARIRenderer ren = anariNewRenderer(d, renderer);
  const float bg[4] = {0.1f, 0.1f, 0.1f, 1.f};
  const float rate = 0.5f;
  anariSetParameter(d, ren, "background", ANARI_FLOAT32_VEC4, bg);
  anariSetParameter(d, ren, "volumeSamplingRate", ANARI_FLOAT32, &rate);
  anariCommitParameters(d, ren);

  ANARIFrame frame = anariNewFrame(d);
  const ANARIDataType colorType = ANARI_UFIXED8_RGBA_SRGB;
  anariSetParameter(d, frame, "size", ANARI_UINT32_VEC2, size);
  anariSetParameter(d, frame, "channel.color", ANARI_DATA_TYPE, &colorType);
  anariSetParameter(d, frame, "renderer", ANARI_RENDERER, &ren);
  anariSetParameter(d, frame, "camera", ANARI_CAMERA, &camera);
  anariSetParameter(d, frame, "world", ANARI_WORLD, &world);
  anariCommitParameters(d, frame);

  float seconds = 0.f;
  for (int i = 0; i < frames; ++i) {
    anariRenderFrame(d, frame);
    anariFrameReady(d, frame, ANARI_WAIT);
    float dur = 0.f;
    anariGetProperty(d, frame, "duration", ANARI_FLOAT32, &dur, sizeof(dur), ANARI_NO_WAIT);
    seconds += dur;
  }

  uint32_t w = 0, h = 0;
  ANARIDataType t = ANARI_UNKNOWN;
  const uint32_t *px = (const uint32_t *)anariMapFrame(d, frame, "channel.color", &w, &h, &t);
  int rc = 1;
  if (px && t == ANARI_UFIXED8_RGBA_SRGB) {
    FILE *f = fopen(argv[2], "wb");
    if (f) {
      fprintf(f, "P6\n%u %u\n255\n", w, h);
      for (uint32_t y = 0; y < h; ++y) /* ANARI's origin is the lower left corner: flip for PPM */
        for (uint32_t x = 0; x < w; ++x) {
          const uint32_t v = px[(size_t)(h - 1 - y) * w + x];
          const unsigned char rgb[3] = {(unsigned char)(v & 0xff), (unsigned char)((v >> 8) & 0xff), (unsigned char)((v >> 16) & 0xff)};
          fwrite(rgb, 1, 3, f);
        }
      fclose(f);
      rc = 0;
    }
  }
  anariUnmapFrame(d, frame, "channel.color");
  printf("%s: %s %ux%ux%u, value range [%g, %g], %d frames of %ux%u with renderer '%s' in %.3f ms (device time)\n", vf.name,
      vf.kind == DVR_IMPORT_NANOVDB ? "nanovdb" : "structuredRegular", vf.dims[0], vf.dims[1], vf.dims[2], vf.valueRange[0],
      vf.valueRange[1], frames, w, h, renderer, seconds * 1e3f);

  anariRelease(d, frame);
  anariRelease(d, ren);
  anariRelease(d, camera);
  anariRelease(d, world);
  anariRelease(d, volume);
  anariRelease(d, field);
  anariRelease(d, d);
  dvr_import_free(&vf);
  return rc;
}
