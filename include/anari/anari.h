/*
 * anari.h — the Khronos ANARI 1.0 C API surface serviced by the B200 DVR device.
 *
 * The ANARI-SDK (libanari frontend, helium) is not available in this environment, so the
 * function signatures, handle types and enum NAMES below are restated from the ANARI 1.0
 * specification; enum numeric values follow the SDK's anari_enums.h as far as they could be
 * recalled and are NOT binary-verified against a real libanari ("ABI").
 * An application written against the SDK's <anari/anari.h> compiles unchanged against this
 * header; the device library (libanari_library_visrtx_b200.so) implements these entry points
 * directly (it is its own frontend) for the object/parameter surface of VisRTX's DVR path.
 */
#ifndef ANARI_B200_ANARI_H
#define ANARI_B200_ANARI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#define ANARI_DEFAULT_VAL(a) = a
#else
#define ANARI_DEFAULT_VAL(a)
#endif

#if defined(__GNUC__)
#define ANARI_INTERFACE __attribute__((visibility("default")))
#else
#define ANARI_INTERFACE
#endif

typedef int ANARIDataType;
typedef int ANARIStatusSeverity;
typedef int ANARIStatusCode;
typedef uint32_t ANARIWaitMask;

/* -------- data types -------- */
#define ANARI_UNKNOWN 0
#define ANARI_DATA_TYPE 100
#define ANARI_STRING 101
#define ANARI_VOID_POINTER 102
#define ANARI_BOOL 103
#define ANARI_STRING_LIST 150
#define ANARI_DATA_TYPE_LIST 151
#define ANARI_PARAMETER_LIST 152
#define ANARI_FUNCTION_POINTER 200
#define ANARI_MEMORY_DELETER 201
#define ANARI_STATUS_CALLBACK 202
#define ANARI_FRAME_COMPLETION_CALLBACK 203
#define ANARI_LIBRARY 500
#define ANARI_DEVICE 501
#define ANARI_OBJECT 502
#define ANARI_ARRAY 503
#define ANARI_ARRAY1D 504
#define ANARI_ARRAY2D 505
#define ANARI_ARRAY3D 506
#define ANARI_CAMERA 507
#define ANARI_FRAME 508
#define ANARI_GEOMETRY 509
#define ANARI_GROUP 510
#define ANARI_INSTANCE 511
#define ANARI_LIGHT 512
#define ANARI_MATERIAL 513
#define ANARI_RENDERER 514
#define ANARI_SURFACE 515
#define ANARI_SAMPLER 516
#define ANARI_SPATIAL_FIELD 517
#define ANARI_VOLUME 518
#define ANARI_WORLD 519
#define ANARI_INT8 1000
#define ANARI_INT8_VEC2 1001
#define ANARI_INT8_VEC3 1002
#define ANARI_INT8_VEC4 1003
#define ANARI_UINT8 1004
#define ANARI_UINT8_VEC2 1005
#define ANARI_UINT8_VEC3 1006
#define ANARI_UINT8_VEC4 1007
#define ANARI_INT16 1008
#define ANARI_UINT16 1012
#define ANARI_INT32 1016
#define ANARI_INT32_VEC2 1017
#define ANARI_INT32_VEC3 1018
#define ANARI_INT32_VEC4 1019
#define ANARI_UINT32 1020
#define ANARI_UINT32_VEC2 1021
#define ANARI_UINT32_VEC3 1022
#define ANARI_UINT32_VEC4 1023
#define ANARI_INT64 1024
#define ANARI_UINT64 1028
#define ANARI_FIXED8 1032
#define ANARI_UFIXED8 1036
#define ANARI_UFIXED8_VEC2 1037
#define ANARI_UFIXED8_VEC3 1038
#define ANARI_UFIXED8_VEC4 1039
#define ANARI_FIXED16 1040
#define ANARI_UFIXED16 1044
#define ANARI_UFIXED16_VEC2 1045
#define ANARI_UFIXED16_VEC3 1046
#define ANARI_UFIXED16_VEC4 1047
#define ANARI_FIXED32 1048
#define ANARI_UFIXED32 1052
#define ANARI_UFIXED32_VEC2 1053
#define ANARI_UFIXED32_VEC3 1054
#define ANARI_UFIXED32_VEC4 1055
#define ANARI_FLOAT16 1064
#define ANARI_FLOAT32 1068
#define ANARI_FLOAT32_VEC2 1069
#define ANARI_FLOAT32_VEC3 1070
#define ANARI_FLOAT32_VEC4 1071
#define ANARI_FLOAT64 1072
#define ANARI_FLOAT64_VEC2 1073
#define ANARI_FLOAT64_VEC3 1074
#define ANARI_FLOAT64_VEC4 1075
#define ANARI_UFIXED8_R_SRGB 2000
#define ANARI_UFIXED8_RA_SRGB 2001
#define ANARI_UFIXED8_RGB_SRGB 2002
#define ANARI_UFIXED8_RGBA_SRGB 2003
#define ANARI_INT32_BOX1 2004
#define ANARI_FLOAT32_BOX1 2008
#define ANARI_FLOAT32_BOX2 2009
#define ANARI_FLOAT32_BOX3 2010
#define ANARI_FLOAT32_BOX4 2011
#define ANARI_FLOAT32_MAT2 2012
#define ANARI_FLOAT32_MAT3 2013
#define ANARI_FLOAT32_MAT4 2014
#define ANARI_FLOAT32_MAT2x3 2015
#define ANARI_FLOAT32_MAT3x4 2016
#define ANARI_FLOAT32_QUAT_IJKW 2017
#define ANARI_UINT64_REGION1 2104
#define ANARI_FLOAT64_BOX1 2208

/* -------- status -------- */
#define ANARI_STATUS_NO_ERROR 0
#define ANARI_STATUS_UNKNOWN_ERROR 1
#define ANARI_STATUS_INVALID_ARGUMENT 2
#define ANARI_STATUS_INVALID_OPERATION 3
#define ANARI_STATUS_OUT_OF_MEMORY 4
#define ANARI_STATUS_UNSUPPORTED_DEVICE 5
#define ANARI_STATUS_VERSION_MISMATCH 6
#define ANARI_SEVERITY_FATAL_ERROR 6000
#define ANARI_SEVERITY_ERROR 6001
#define ANARI_SEVERITY_WARNING 6002
#define ANARI_SEVERITY_PERFORMANCE_WARNING 6003
#define ANARI_SEVERITY_INFO 6004
#define ANARI_SEVERITY_DEBUG 6005
#define ANARI_NO_WAIT 0
#define ANARI_WAIT 1

/* -------- handles -------- */
typedef struct _ANARILibrary *ANARILibrary;
typedef struct _ANARIObject *ANARIObject;
typedef ANARIObject ANARIDevice;
typedef ANARIObject ANARIArray;
typedef ANARIObject ANARIArray1D;
typedef ANARIObject ANARIArray2D;
typedef ANARIObject ANARIArray3D;
typedef ANARIObject ANARICamera;
typedef ANARIObject ANARIFrame;
typedef ANARIObject ANARIGeometry;
typedef ANARIObject ANARIGroup;
typedef ANARIObject ANARIInstance;
typedef ANARIObject ANARILight;
typedef ANARIObject ANARIMaterial;
typedef ANARIObject ANARIRenderer;
typedef ANARIObject ANARISampler;
typedef ANARIObject ANARISurface;
typedef ANARIObject ANARISpatialField;
typedef ANARIObject ANARIVolume;
typedef ANARIObject ANARIWorld;

typedef struct
{
  const char *name;
  ANARIDataType type;
} ANARIParameter;

typedef void (*ANARIMemoryDeleter)(const void *userPtr, const void *appMemory);
typedef void (*ANARIStatusCallback)(const void *userPtr, ANARIDevice device, ANARIObject source,
    ANARIDataType sourceType, ANARIStatusSeverity severity, ANARIStatusCode code, const char *message);
typedef void (*ANARIFrameCompletionCallback)(const void *userPtr, ANARIDevice device, ANARIFrame frame);

/* -------- library / device -------- */
ANARI_INTERFACE ANARILibrary anariLoadLibrary(const char *name, ANARIStatusCallback statusCallback ANARI_DEFAULT_VAL(0),
    const void *statusCallbackUserData ANARI_DEFAULT_VAL(0));
ANARI_INTERFACE void anariUnloadLibrary(ANARILibrary library);
ANARI_INTERFACE void anariLoadModule(ANARILibrary library, const char *name);
ANARI_INTERFACE void anariUnloadModule(ANARILibrary library, const char *name);
ANARI_INTERFACE const char **anariGetDeviceSubtypes(ANARILibrary library);
ANARI_INTERFACE const char **anariGetDeviceExtensions(ANARILibrary library, const char *deviceSubtype);
ANARI_INTERFACE ANARIDevice anariNewDevice(ANARILibrary library, const char *type ANARI_DEFAULT_VAL("default"));

/* -------- arrays -------- */
ANARI_INTERFACE ANARIArray1D anariNewArray1D(ANARIDevice device, const void *appMemory, ANARIMemoryDeleter deleter,
    const void *userData, ANARIDataType dataType, uint64_t numItems1);
ANARI_INTERFACE ANARIArray2D anariNewArray2D(ANARIDevice device, const void *appMemory, ANARIMemoryDeleter deleter,
    const void *userData, ANARIDataType dataType, uint64_t numItems1, uint64_t numItems2);
ANARI_INTERFACE ANARIArray3D anariNewArray3D(ANARIDevice device, const void *appMemory, ANARIMemoryDeleter deleter,
    const void *userData, ANARIDataType dataType, uint64_t numItems1, uint64_t numItems2, uint64_t numItems3);
ANARI_INTERFACE void *anariMapArray(ANARIDevice device, ANARIArray array);
ANARI_INTERFACE void anariUnmapArray(ANARIDevice device, ANARIArray array);

/* -------- objects -------- */
ANARI_INTERFACE ANARILight anariNewLight(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARICamera anariNewCamera(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARIGeometry anariNewGeometry(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARISpatialField anariNewSpatialField(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARIVolume anariNewVolume(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARISurface anariNewSurface(ANARIDevice device);
ANARI_INTERFACE ANARIMaterial anariNewMaterial(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARISampler anariNewSampler(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARIGroup anariNewGroup(ANARIDevice device);
ANARI_INTERFACE ANARIInstance anariNewInstance(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARIWorld anariNewWorld(ANARIDevice device);
ANARI_INTERFACE ANARIObject anariNewObject(ANARIDevice device, const char *objectType, const char *type);
ANARI_INTERFACE ANARIRenderer anariNewRenderer(ANARIDevice device, const char *type);
ANARI_INTERFACE ANARIFrame anariNewFrame(ANARIDevice device);

/* -------- parameters / lifetime -------- */
ANARI_INTERFACE void anariSetParameter(ANARIDevice device, ANARIObject object, const char *name, ANARIDataType dataType,
    const void *mem);
ANARI_INTERFACE void anariUnsetParameter(ANARIDevice device, ANARIObject object, const char *name);
ANARI_INTERFACE void anariUnsetAllParameters(ANARIDevice device, ANARIObject object);
ANARI_INTERFACE void *anariMapParameterArray1D(ANARIDevice device, ANARIObject object, const char *name,
    ANARIDataType dataType, uint64_t numElements1, uint64_t *elementStride);
ANARI_INTERFACE void *anariMapParameterArray2D(ANARIDevice device, ANARIObject object, const char *name,
    ANARIDataType dataType, uint64_t numElements1, uint64_t numElements2, uint64_t *elementStride);
ANARI_INTERFACE void *anariMapParameterArray3D(ANARIDevice device, ANARIObject object, const char *name,
    ANARIDataType dataType, uint64_t numElements1, uint64_t numElements2, uint64_t numElements3,
    uint64_t *elementStride);
ANARI_INTERFACE void anariUnmapParameterArray(ANARIDevice device, ANARIObject object, const char *name);
ANARI_INTERFACE void anariCommitParameters(ANARIDevice device, ANARIObject object);
ANARI_INTERFACE void anariRelease(ANARIDevice device, ANARIObject object);
ANARI_INTERFACE void anariRetain(ANARIDevice device, ANARIObject object);

/* -------- introspection / properties -------- */
ANARI_INTERFACE const char **anariGetObjectSubtypes(ANARIDevice device, ANARIDataType objectType);
ANARI_INTERFACE const void *anariGetObjectInfo(ANARIDevice device, ANARIDataType objectType, const char *objectSubtype,
    const char *infoName, ANARIDataType infoType);
ANARI_INTERFACE const void *anariGetParameterInfo(ANARIDevice device, ANARIDataType objectType,
    const char *objectSubtype, const char *parameterName, ANARIDataType parameterType, const char *infoName,
    ANARIDataType infoType);
ANARI_INTERFACE int anariGetProperty(ANARIDevice device, ANARIObject object, const char *name, ANARIDataType type,
    void *mem, uint64_t size, ANARIWaitMask mask);

/* -------- frames -------- */
ANARI_INTERFACE const void *anariMapFrame(ANARIDevice device, ANARIFrame frame, const char *channel, uint32_t *width,
    uint32_t *height, ANARIDataType *pixelType);
ANARI_INTERFACE void anariUnmapFrame(ANARIDevice device, ANARIFrame frame, const char *channel);
ANARI_INTERFACE void anariRenderFrame(ANARIDevice device, ANARIFrame frame);
ANARI_INTERFACE int anariFrameReady(ANARIDevice device, ANARIFrame frame, ANARIWaitMask mask);
ANARI_INTERFACE void anariDiscardFrame(ANARIDevice device, ANARIFrame frame);

#ifdef __cplusplus
}
#endif
#endif
