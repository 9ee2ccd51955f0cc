/* dvr_import.h — C-ABI of the volume-file importers that feed the DVR path.
 *
 * Host-only helpers (no CUDA): they turn a volume file into exactly the parameters the `structuredRegular` /
 * `nanovdb` spatial field takes, the way the reference's TSD importers do:
 *
 *   dvr_import_raw     tsd/src/tsd/authoring/importers/import_RAW.cpp:12-86     (dims and type from the file name)
 *   dvr_import_mhd     tsd/src/tsd/authoring/importers/import_MHD.cpp:18-115    (MetaImage header + raw data file)
 *   dvr_import_vti     tsd/src/tsd/authoring/importers/import_VTI.cpp:57-113    (VTK XML ImageData; the reference
 *                      calls vtkXMLImageDataReader — VTK is a third-party dependency absent from the reference
 *                      tree, so the published VTK XML format is restated here: ascii / inline binary / appended,
 *                      raw or base64, UInt32 / UInt64 headers, optional vtkZLibDataCompressor)
 *   dvr_import_nvdb    tsd/src/tsd/authoring/importers/import_NVDB.cpp:17-115   (nanovdb::io::readGrid: segment
 *                      files with codec NONE / ZIP and raw grid buffers; min/max from the root, computed when the
 *                      grid carries none — nanovdb::tools::updateGridStats(StatsMode::MinMax))
 *   dvr_import_volume  tsd/src/tsd/authoring/importers/import_volume.cpp:12-66  (dispatch on the extension)
 *   dvr_compute_scalar_range  tsd/src/tsd/algorithms/computeScalarRange.cpp:12-66 (normalised min/max of an array)
 *
 * Error behaviour follows the importers: a file that cannot be parsed or read yields an error code and a message
 * (dvr_import_last_error), never an exception or a partially filled result.
 */
#ifndef DVR_IMPORT_H
#define DVR_IMPORT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum
{
  DVR_IMPORT_OK = 0,
  DVR_IMPORT_ERR_ARGUMENT = -1,   /* null pointer, unknown extension */
  DVR_IMPORT_ERR_IO = -2,         /* cannot open / short read */
  DVR_IMPORT_ERR_FORMAT = -3,     /* file does not parse */
  DVR_IMPORT_ERR_UNSUPPORTED = -4 /* valid file, feature outside the importer (e.g. BLOSC codec, 32-bit ints) */
};

enum
{
  DVR_IMPORT_STRUCTURED = 0, /* data = dims[0]*dims[1]*dims[2] voxels, x fastest: the `data` ARRAY3D of structuredRegular */
  DVR_IMPORT_NANOVDB = 1     /* data = one serialized NanoVDB grid: the `data` UINT8 ARRAY1D of a nanovdb field */
};

typedef struct DvrVolumeFile
{
  int32_t kind;         /* DVR_IMPORT_STRUCTURED | DVR_IMPORT_NANOVDB */
  int32_t dataType;     /* DvrDataType of include/dvr_b200.h (structured) ; -1 for nanovdb */
  uint32_t dims[3];     /* structured: voxel counts; nanovdb: extent of the index bounding box */
  float origin[3];      /* field `origin` (VTI: the file's; RAW/MHD: 0 like the reference, which never sets it) */
  float spacing[3];     /* field `spacing` (VTI: the file's; RAW/MHD: 1 — import_MHD parses ElementSpacing but does
                           not apply it; the parsed values are reported in headerSpacing) */
  double headerSpacing[3];
  float valueRange[2];  /* what import_volume puts in the volume's `valueRange`: computeScalarRange of the data
                           (normalised for fixed-point types) or the NanoVDB root min/max; {0,1} when unknown */
  int32_t hasValueRange;
  int32_t _pad;
  void *data;           /* malloc'ed; release with dvr_import_free */
  uint64_t bytes;
  char name[256];       /* object name the importer assigns (file name) */
} DvrVolumeFile;

const char *dvr_import_last_error(void);
int dvr_import_raw(const char *path, DvrVolumeFile *out);
int dvr_import_mhd(const char *path, DvrVolumeFile *out);
int dvr_import_vti(const char *path, DvrVolumeFile *out);
int dvr_import_nvdb(const char *path, DvrVolumeFile *out);
int dvr_import_volume(const char *path, DvrVolumeFile *out);
void dvr_import_free(DvrVolumeFile *f);

/* computeScalarRange: min/max of n elements of a host array, converted like ANARITypeProperties<T>::toFloat4
 * (UFIXED8 v/255, FIXED8 max(v/127,-1), UFIXED16 v/65535, FIXED16 max(v/32767,-1), FLOAT32, FLOAT64).
 * out = {FLT_MAX,-FLT_MAX} and DVR_IMPORT_ERR_UNSUPPORTED for other types, like the reference's warning path. */
int dvr_compute_scalar_range(const void *hostData, int32_t dataType, uint64_t n, float out[2]);

#ifdef __cplusplus
}
#endif
#endif
