/*
 * visrtx_b200.h — direct device construction and extension queries of the B200 DVR device.
 * Mirrors devices/rtx/include/anari/ext/visrtx/visrtx.h:46-73 of the reference
 * (makeVisRTXDevice, VisRTXExtensions, visrtxGetObjectExtensions, visrtxGetInstanceExtensions).
 */
#ifndef ANARI_EXT_VISRTX_B200_H
#define ANARI_EXT_VISRTX_B200_H
#include <anari/anari.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same name as the reference so that `ANARIDevice d = makeVisRTXDevice(cb, ptr);` applications relink */
ANARI_INTERFACE ANARIDevice makeVisRTXDevice(ANARIStatusCallback defaultCallback ANARI_DEFAULT_VAL(0),
    const void *userPtr ANARI_DEFAULT_VAL(0));

typedef struct
{
  int VISRTX_ARRAY_CUDA;
  int VISRTX_CUDA_OUTPUT_BUFFERS;
  int VISRTX_INSTANCE_ATTRIBUTES;
  int VISRTX_MATERIAL_MDL;
  int VISRTX_SPATIAL_FIELD_NANOVDB;
  int VISRTX_TRIANGLE_BACK_FACE_CULLING;
  int VISRTX_TRIANGLE_FACE_VARYING_ATTRIBUTES;
} VisRTXExtensions;

ANARI_INTERFACE int visrtxGetObjectExtensions(VisRTXExtensions *extensions, ANARIDevice device,
    ANARIDataType objectType, const char *objectSubtype);
ANARI_INTERFACE int visrtxGetInstanceExtensions(VisRTXExtensions *extensions, ANARIDevice device, ANARIObject object);

#ifdef __cplusplus
}
#endif
#endif
