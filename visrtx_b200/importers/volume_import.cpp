// volume_import.cpp — host-side volume-file importers behind include/dvr_import.h.
//
// Each function restates one of the reference's TSD importers (file:line in the header) and produces the raw
// parameters of the spatial field it would create.  Pure C++17 + zlib; no CUDA, no ANARI.
#include "dvr_import.h"
#include "dvr_b200.h"
#include "dvr_nvdb_validate.h"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// Path / token helpers with the behaviour of the TSD importers' (importer_common.cpp:41-64): the base name is
// everything after the last '/', empty when the path has no directory part; the extension includes its dot.
std::string fileOf(const std::string &path)
{
  const char *begin = path.c_str(), *slash = std::strrchr(begin, '/');
  return slash ? std::string(slash + 1) : std::string();
}
std::string extensionOf(const std::string &path)
{
  const char *begin = path.c_str(), *dot = std::strrchr(begin, '.');
  return dot ? std::string(dot) : std::string();
}
// getline semantics: consecutive delimiters yield empty tokens, a trailing delimiter yields none
std::vector<std::string> splitString(const std::string &text, char delim)
{
  std::vector<std::string> tokens;
  size_t from = 0;
  while (from < text.size()) {
    size_t to = text.find(delim, from);
    if (to == std::string::npos)
      to = text.size();
    tokens.emplace_back(text, from, to - from);
    from = to + 1;
  }
  return tokens;
}

size_t sizeOfType(int t)
{
  switch (t) {
  case DVR_UFIXED8:
  case DVR_FIXED8: return 1;
  case DVR_UFIXED16:
  case DVR_FIXED16:
  case DVR_FLOAT16: return 2;
  case DVR_FLOAT32: return 4;
  case DVR_FLOAT64: return 8;
  default: return 0;
  }
}

void initOut(DvrVolumeFile *o)
{
  std::memset(o, 0, sizeof(*o));
  o->spacing[0] = o->spacing[1] = o->spacing[2] = 1.f;
  o->headerSpacing[0] = o->headerSpacing[1] = o->headerSpacing[2] = 1.0;
  o->valueRange[0] = 0.f;
  o->valueRange[1] = 1.f;
}

void setName(DvrVolumeFile *o, const std::string &n)
{
  std::snprintf(o->name, sizeof(o->name), "%s", n.c_str());
}

template <typename T, typename F>
void rangeOf(const void *p, uint64_t n, F toFloat, float out[2])
{
  const T *b = static_cast<const T *>(p);
  const auto mm = std::minmax_element(b, b + n); // computeScalarRangeImpl.hpp:42-46: extrema first, then convert
  out[0] = toFloat(*mm.first);
  out[1] = toFloat(*mm.second);
}

int scalarRange(const void *data, int32_t type, uint64_t n, float out[2])
{
  out[0] = FLT_MAX;
  out[1] = -FLT_MAX;
  if (!data || n == 0)
    return fail(DVR_IMPORT_ERR_ARGUMENT, "computeScalarRange: empty array");
  switch (type) {
  case DVR_UFIXED8: rangeOf<uint8_t>(data, n, [](uint8_t v) { return (float)v / 255.f; }, out); break;
  case DVR_FIXED8: rangeOf<int8_t>(data, n, [](int8_t v) { return std::max((float)v / 127.f, -1.f); }, out); break;
  case DVR_UFIXED16: rangeOf<uint16_t>(data, n, [](uint16_t v) { return (float)v / 65535.f; }, out); break;
  case DVR_FIXED16: rangeOf<int16_t>(data, n, [](int16_t v) { return std::max((float)v / 32767.f, -1.f); }, out); break;
  case DVR_FLOAT32: rangeOf<float>(data, n, [](float v) { return v; }, out); break;
  case DVR_FLOAT64: rangeOf<double>(data, n, [](double v) { return (float)v; }, out); break;
  default:
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "computeScalarRange() called on an array with incompatible element type %d",
        (int)type);
  }
  return DVR_IMPORT_OK;
}

int finishStructured(DvrVolumeFile *o)
{
  const uint64_t n = (uint64_t)o->dims[0] * o->dims[1] * o->dims[2];
  float r[2];
  if (scalarRange(o->data, o->dataType, n, r) == DVR_IMPORT_OK) { // SpatialField::computeValueRange, SpatialField.cpp:81-83
    o->valueRange[0] = r[0];
    o->valueRange[1] = r[1];
    o->hasValueRange = 1;
  }
  return DVR_IMPORT_OK;
}

int readWhole(const std::string &path, uint64_t bytes, void **out, const char *who)
{
  FILE *fh = std::fopen(path.c_str(), "rb");
  if (!fh)
    return fail(DVR_IMPORT_ERR_IO, "[%s] unable to open RAW file: '%s'", who, path.c_str());
  void *buf = std::malloc(bytes ? bytes : 1);
  if (!buf) {
    std::fclose(fh);
    return fail(DVR_IMPORT_ERR_IO, "[%s] out of memory (%llu bytes)", who, (unsigned long long)bytes);
  }
  const size_t got = std::fread(buf, 1, bytes, fh);
  std::fclose(fh);
  if (got != bytes) { // the reference's fread(ptr, size, 1, f) fails the same way on a short file
    std::free(buf);
    return fail(DVR_IMPORT_ERR_IO, "[%s] unable to open RAW file: '%s' (%llu of %llu bytes)", who, path.c_str(),
        (unsigned long long)got, (unsigned long long)bytes);
  }
  *out = buf;
  return DVR_IMPORT_OK;
}

// ---- RAW (import_RAW.cpp:12-86) -------------------------------------------------------------------------
int importRaw(const char *filepath, DvrVolumeFile *o)
{
  const std::string file = fileOf(filepath);
  if (file.empty())
    return fail(DVR_IMPORT_ERR_ARGUMENT, "[import_RAW] no file name in path '%s'", filepath);
  int dimX = 0, dimY = 0, dimZ = 0;
  int type = -1;
  bool unsupported32 = false;
  for (const auto &str : splitString(file, '_')) {
    int x = 0, y = 0, z = 0;
    if (std::sscanf(str.c_str(), "%ix%ix%i", &x, &y, &z) == 3) {
      dimX = x;
      dimY = y;
      dimZ = z;
    }
    // "int%i" and "uint%i" both select the UNSIGNED normalised types (import_RAW.cpp:32-50)
    int bits = 0;
    for (const char *fmt : {"int%i", "uint%i"})
      if (std::sscanf(str.c_str(), fmt, &bits) == 1) {
        if (bits == 8)
          type = DVR_UFIXED8;
        else if (bits == 16)
          type = DVR_UFIXED16;
        else if (bits == 32)
          unsupported32 = true; // ANARI_UFIXED32: a type the structuredRegular field itself rejects
      }
    if (dimX && dimY && dimZ && type != -1)
      break;
  }
  if (type == -1 && unsupported32)
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_RAW] 32-bit integer voxels are not a structuredRegular data type: '%s'",
        file.c_str());
  if (type == -1)
    type = DVR_FLOAT32;
  if (!(dimX && dimY && dimZ))
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_RAW] unable to parse info from RAW file: '%s'", file.c_str());
  if (dimX < 0 || dimY < 0 || dimZ < 0)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_RAW] negative dimensions in RAW file name: '%s'", file.c_str());
  o->kind = DVR_IMPORT_STRUCTURED;
  o->dataType = type;
  o->dims[0] = (uint32_t)dimX;
  o->dims[1] = (uint32_t)dimY;
  o->dims[2] = (uint32_t)dimZ;
  setName(o, file);
  o->bytes = (uint64_t)dimX * (uint64_t)dimY * (uint64_t)dimZ * sizeOfType(type);
  const int rc = readWhole(filepath, o->bytes, &o->data, "import_RAW");
  return rc == DVR_IMPORT_OK ? finishStructured(o) : rc;
}

// ---- MHD (import_MHD.cpp:18-115) ------------------------------------------------------------------------
std::string trim(std::string s)
{
  const size_t a = s.find_first_not_of(" \t\r");
  if (a == std::string::npos)
    return "";
  s.erase(0, a);
  s.erase(s.find_last_not_of(" \t\r") + 1);
  return s;
}

int importMhd(const char *filepath, DvrVolumeFile *o)
{
  std::ifstream file(filepath);
  if (!file.is_open())
    return fail(DVR_IMPORT_ERR_IO, "[import_MHD] unable to open header '%s'", filepath);
  unsigned dims[3] = {0, 0, 0};
  double spacing[3] = {1.0, 1.0, 1.0};
  int type = -1;
  std::string dataFile, typeName;
  std::string line;
  while (std::getline(file, line)) {
    const size_t delim = line.find('=');
    if (delim == std::string::npos)
      continue;
    const std::string key = trim(line.substr(0, delim)), value = trim(line.substr(delim + 1));
    if (key == "DimSize")
      std::sscanf(value.c_str(), "%u %u %u", &dims[0], &dims[1], &dims[2]);
    else if (key == "ElementSpacing")
      std::sscanf(value.c_str(), "%lf %lf %lf", &spacing[0], &spacing[1], &spacing[2]);
    else if (key == "ElementType") {
      typeName = value;
      // the reference maps exactly these three (MET_SHORT to the UNSIGNED 16-bit type), import_MHD.cpp:60-67;
      // the others below are this importer's additions for types the field supports
      if (value == "MET_UCHAR")
        type = DVR_UFIXED8;
      else if (value == "MET_SHORT")
        type = DVR_UFIXED16;
      else if (value == "MET_FLOAT")
        type = DVR_FLOAT32;
      else if (value == "MET_USHORT")
        type = DVR_UFIXED16;
      else if (value == "MET_CHAR")
        type = DVR_FIXED8;
      else if (value == "MET_DOUBLE")
        type = DVR_FLOAT64;
    } else if (key == "ElementDataFile")
      dataFile = value;
    // BinaryData / BinaryDataByteOrderMSB are parsed and ignored by the reference (import_MHD.cpp:71-75)
  }
  if (type == -1)
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_MHD] unsupported ElementType '%s' in '%s'", typeName.c_str(), filepath);
  if (!(dims[0] && dims[1] && dims[2]) || dataFile.empty())
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_MHD] header '%s' lacks DimSize / ElementDataFile", filepath);
  const std::string p(filepath);
  const size_t slash = p.find_last_of('/');
  const std::string dataPath = (slash == std::string::npos ? std::string() : p.substr(0, slash)) + "/" + dataFile;
  o->kind = DVR_IMPORT_STRUCTURED;
  o->dataType = type;
  for (int i = 0; i < 3; ++i) {
    o->dims[i] = dims[i];
    o->headerSpacing[i] = spacing[i];
  }
  setName(o, dataPath);
  o->bytes = (uint64_t)dims[0] * dims[1] * dims[2] * sizeOfType(type);
  const int rc = readWhole(slash == std::string::npos ? dataFile : dataPath, o->bytes, &o->data, "import_MHD");
  return rc == DVR_IMPORT_OK ? finishStructured(o) : rc;
}

// ---- VTI: VTK XML ImageData (import_VTI.cpp:57-113 through vtkXMLImageDataReader) ------------------------
struct XmlTag
{
  std::string name;
  std::map<std::string, std::string> attr;
  size_t begin = 0, end = 0; // [begin,end) of the tag text; content starts at end
  bool selfClosing = false;
};

// next tag at or after pos (skips comments / declarations); false at the end of text
bool nextTag(const std::string &t, size_t pos, size_t limit, XmlTag &tag)
{
  while (true) {
    const size_t lt = t.find('<', pos);
    if (lt == std::string::npos || lt >= limit)
      return false;
    if (t.compare(lt, 4, "<!--") == 0) {
      const size_t e = t.find("-->", lt);
      if (e == std::string::npos)
        return false;
      pos = e + 3;
      continue;
    }
    if (lt + 1 < t.size() && (t[lt + 1] == '?' || t[lt + 1] == '!' || t[lt + 1] == '/')) {
      const size_t e = t.find('>', lt);
      if (e == std::string::npos)
        return false;
      pos = e + 1;
      continue;
    }
    const size_t gt = t.find('>', lt);
    if (gt == std::string::npos)
      return false;
    tag = XmlTag();
    tag.begin = lt;
    tag.end = gt + 1;
    std::string body = t.substr(lt + 1, gt - lt - 1);
    if (!body.empty() && body.back() == '/') {
      tag.selfClosing = true;
      body.pop_back();
    }
    size_t i = 0;
    while (i < body.size() && !std::isspace((unsigned char)body[i]))
      ++i;
    tag.name = body.substr(0, i);
    while (i < body.size()) {
      while (i < body.size() && std::isspace((unsigned char)body[i]))
        ++i;
      const size_t eq = body.find('=', i);
      if (eq == std::string::npos)
        break;
      const std::string key = trim(body.substr(i, eq - i));
      size_t q = eq + 1;
      while (q < body.size() && std::isspace((unsigned char)body[q]))
        ++q;
      if (q >= body.size() || (body[q] != '"' && body[q] != '\''))
        break;
      const size_t qe = body.find(body[q], q + 1);
      if (qe == std::string::npos)
        break;
      tag.attr[key] = body.substr(q + 1, qe - q - 1);
      i = qe + 1;
    }
    return true;
  }
}

bool base64Decode(const char *s, size_t n, std::vector<uint8_t> &out, size_t maxBytes = SIZE_MAX, size_t *consumed = nullptr)
{
  static int8_t lut[256];
  static bool init = false;
  if (!init) {
    std::memset(lut, -1, sizeof(lut));
    const char *abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i)
      lut[(unsigned char)abc[i]] = (int8_t)i;
    init = true;
  }
  uint32_t acc = 0;
  int bits = 0, quad = 0;
  size_t i = 0;
  for (; i < n && out.size() < maxBytes; ++i) {
    const unsigned char c = (unsigned char)s[i];
    if (std::isspace(c))
      continue;
    if (c == '=') { // padding ends one base64 block
      ++quad;
      if (quad % 4 == 0) {
        bits = 0;
        acc = 0;
        ++i;
        if (maxBytes != SIZE_MAX)
          break;
      }
      continue;
    }
    if (lut[c] < 0)
      break;
    acc = (acc << 6) | (uint32_t)lut[c];
    bits += 6;
    ++quad;
    if (bits >= 8) {
      bits -= 8;
      out.push_back((uint8_t)((acc >> bits) & 0xff));
    }
    if (quad % 4 == 0 && maxBytes != SIZE_MAX && out.size() >= maxBytes) {
      ++i;
      break;
    }
  }
  if (consumed)
    *consumed = i;
  return true;
}

struct VtkType
{
  const char *name;
  int dvrType; // -1: readable but not a field type
  size_t size;
  char kind; // 'f' float, 'i' signed, 'u' unsigned
};
const VtkType kVtkTypes[] = {{"Float32", DVR_FLOAT32, 4, 'f'}, {"Float64", DVR_FLOAT64, 8, 'f'}, {"Int8", DVR_FIXED8, 1, 'i'},
    {"UInt8", DVR_UFIXED8, 1, 'u'}, {"Int16", DVR_FIXED16, 2, 'i'}, {"UInt16", DVR_UFIXED16, 2, 'u'},
    {"Int32", -1, 4, 'i'}, {"UInt32", -1, 4, 'u'}, {"Int64", -1, 8, 'i'}, {"UInt64", -1, 8, 'u'}};

uint64_t rdHeader(const uint8_t *p, size_t hsz)
{
  if (hsz == 8) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;
  }
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}

// One binary data block of the VTK XML format -> `want` raw bytes.  `src` either holds raw bytes (appended raw)
// or base64 text.  Uncompressed: [nbytes][data].  vtkZLibDataCompressor: [nblocks][blocksize][lastsize][c_0..c_n-1]
// followed by the zlib streams; in base64 mode the header and the payload are encoded separately.
int decodeBlock(const char *src, size_t avail, bool isBase64, bool zlibCompressed, size_t hsz, size_t want,
    std::vector<uint8_t> &out)
{
  std::vector<uint8_t> head, payload;
  const uint8_t *raw = reinterpret_cast<const uint8_t *>(src);
  size_t pos = 0;
  auto take = [&](size_t n, std::vector<uint8_t> &dst) -> bool {
    dst.clear();
    if (isBase64) {
      size_t used = 0;
      base64Decode(src + pos, avail - pos, dst, n, &used);
      pos += used;
      if (dst.size() < n)
        return false;
      dst.resize(n);
      return true;
    }
    if (pos > avail || n > avail - pos) // (pos + n could wrap for a header-supplied n)
      return false;
    dst.assign(raw + pos, raw + pos + n);
    pos += n;
    return true;
  };
  if (!zlibCompressed) {
    if (isBase64) { // header and data are one base64 stream
      std::vector<uint8_t> all;
      base64Decode(src, avail, all, hsz + want);
      if (want > SIZE_MAX - hsz || all.size() < hsz + want)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated base64 data block");
      if (rdHeader(all.data(), hsz) < want)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] data block is smaller than the array it should hold");
      out.assign(all.begin() + hsz, all.begin() + hsz + want);
      return DVR_IMPORT_OK;
    }
    if (!take(hsz, head) || rdHeader(head.data(), hsz) < want || !take(want, out))
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated appended data block");
    return DVR_IMPORT_OK;
  }
  if (!take(3 * hsz, head))
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated compression header");
  const uint64_t nblocks = rdHeader(head.data(), hsz), blockSize = rdHeader(head.data() + hsz, hsz),
                 lastSize = rdHeader(head.data() + 2 * hsz, hsz);
  if (nblocks == 0) {
    out.clear();
    return want == 0 ? DVR_IMPORT_OK : fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] empty compressed block");
  }
  if (nblocks > (1u << 24))
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] implausible compressed block count");
  std::vector<uint8_t> sizes;
  if (isBase64) { // the size table continues the header's base64 stream
    pos = 0;
    if (!take((3 + nblocks) * hsz, head))
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated compression header");
    sizes.assign(head.begin() + 3 * hsz, head.end());
  } else if (!take(nblocks * hsz, sizes))
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated compression header");
  // Every quantity below comes from the file: sizes are bounded by what the array can need (`want` plus one block of
  // slack) and summed with overflow checks, so neither `total` nor the compressed sizes can wrap into a small
  // allocation that uncompress() would then overrun.
  if (blockSize == 0 || blockSize > (uint64_t)1 << 40 || lastSize > blockSize)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] implausible compression block size");
  if ((nblocks - 1) > (UINT64_MAX - blockSize) / blockSize)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] compressed block sizes overflow");
  const uint64_t total = (nblocks - 1) * blockSize + (lastSize ? lastSize : blockSize);
  if (total < want)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] data block is smaller than the array it should hold");
  if (total - want > blockSize) // total >= want here; a well-formed stream inflates to exactly the array
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] data block is larger than the array it should hold");
  uint64_t compressedTotal = 0;
  for (uint64_t b = 0; b < nblocks; ++b) {
    const uint64_t csz = rdHeader(sizes.data() + b * hsz, hsz);
    if (csz > avail || compressedTotal > (uint64_t)avail - csz) // cannot exceed the bytes that are there
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated compressed payload");
    compressedTotal += csz;
  }
  if (!take((size_t)compressedTotal, payload))
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] truncated compressed payload");
  out.resize((size_t)total);
  size_t in = 0, outPos = 0;
  for (uint64_t b = 0; b < nblocks; ++b) {
    const uint64_t csz = rdHeader(sizes.data() + b * hsz, hsz);
    if (csz > payload.size() - in)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] compressed block %llu exceeds the payload", (unsigned long long)b);
    const uint64_t blockOut = b + 1 == nblocks && lastSize ? lastSize : blockSize;
    uLongf dsz = (uLongf)std::min<uint64_t>(blockOut, out.size() - outPos); // never past the buffer
    if (uncompress(out.data() + outPos, &dsz, payload.data() + in, (uLong)csz) != Z_OK)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] zlib error in block %llu", (unsigned long long)b);
    in += (size_t)csz;
    outPos += dsz;
  }
  if (outPos < want)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] compressed blocks inflate to fewer bytes than the array needs");
  out.resize(want);
  return DVR_IMPORT_OK;
}

void byteSwap(uint8_t *p, size_t n, size_t esz)
{
  for (size_t i = 0; i + esz <= n * esz; i += esz)
    std::reverse(p + i, p + i + esz);
}

int importVti(const char *filepath, DvrVolumeFile *o)
{
  std::ifstream is(filepath, std::ios::binary);
  if (!is.is_open())
    return fail(DVR_IMPORT_ERR_IO, "[import_VTI] failed to load .vti file '%s'", filepath);
  std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());

  // the raw appended section must not be scanned as XML
  size_t xmlLimit = text.size(), appendedData = std::string::npos;
  bool appendedBase64 = false;
  XmlTag tag;
  {
    const size_t ap = text.find("<AppendedData");
    if (ap != std::string::npos) {
      XmlTag at;
      if (!nextTag(text, ap, text.size(), at))
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] malformed AppendedData element in '%s'", filepath);
      appendedBase64 = at.attr["encoding"] == "base64";
      const size_t us = text.find('_', at.end);
      if (us == std::string::npos)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] AppendedData without the '_' marker in '%s'", filepath);
      appendedData = us + 1;
      xmlLimit = ap;
    }
  }
  size_t pos = 0;
  XmlTag vtkFile, image, piece;
  bool haveFile = false, haveImage = false;
  std::vector<XmlTag> pointArrays;
  bool inPointData = false;
  while (nextTag(text, pos, xmlLimit, tag)) {
    pos = tag.end;
    if (tag.name == "VTKFile") {
      vtkFile = tag;
      haveFile = true;
    } else if (tag.name == "ImageData") {
      image = tag;
      haveImage = true;
    } else if (tag.name == "Piece")
      piece = tag;
    else if (tag.name == "PointData")
      inPointData = !tag.selfClosing;
    else if (tag.name == "CellData")
      inPointData = false;
    else if (tag.name == "DataArray" && inPointData)
      pointArrays.push_back(tag);
  }
  if (!haveFile || !haveImage || vtkFile.attr["type"] != "ImageData")
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] failed to load .vti file '%s' (not a VTK ImageData file)", filepath);
  const bool bigEndian = vtkFile.attr["byte_order"] == "BigEndian";
  const size_t hsz = vtkFile.attr["header_type"] == "UInt64" ? 8 : 4;
  const std::string compressor = vtkFile.attr["compressor"];
  if (!compressor.empty() && compressor != "vtkZLibDataCompressor")
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_VTI] compressor '%s' is not supported", compressor.c_str());
  if (bigEndian && (!compressor.empty() || hsz != 4))
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_VTI] big-endian files are read only uncompressed with UInt32 headers");

  int ext[6] = {0, -1, 0, -1, 0, -1};
  double spacing[3] = {1.0, 1.0, 1.0}, origin[3] = {0.0, 0.0, 0.0};
  if (std::sscanf(image.attr["WholeExtent"].c_str(), "%d %d %d %d %d %d", &ext[0], &ext[1], &ext[2], &ext[3], &ext[4], &ext[5]) != 6)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] missing WholeExtent in '%s'", filepath);
  if (image.attr.count("Spacing"))
    std::sscanf(image.attr["Spacing"].c_str(), "%lf %lf %lf", &spacing[0], &spacing[1], &spacing[2]);
  if (image.attr.count("Origin"))
    std::sscanf(image.attr["Origin"].c_str(), "%lf %lf %lf", &origin[0], &origin[1], &origin[2]);
  if (piece.attr.count("Extent")) {
    int pe[6];
    if (std::sscanf(piece.attr["Extent"].c_str(), "%d %d %d %d %d %d", &pe[0], &pe[1], &pe[2], &pe[3], &pe[4], &pe[5]) == 6
        && std::memcmp(pe, ext, sizeof(pe)) != 0)
      return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_VTI] multi-piece image data is not supported");
  }
  const int64_t dims[3] = {(int64_t)ext[1] - ext[0] + 1, (int64_t)ext[3] - ext[2] + 1, (int64_t)ext[5] - ext[4] + 1};
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] empty extent in '%s'", filepath);
  const uint64_t nPoints = (uint64_t)dims[0] * dims[1] * dims[2];

  // first single-component point-data array (import_VTI.cpp:93-109: multi-component arrays are skipped)
  const XmlTag *arr = nullptr;
  for (const auto &a : pointArrays) {
    auto it = a.attr.find("NumberOfComponents");
    if (it != a.attr.end() && std::atoi(it->second.c_str()) > 1)
      continue;
    arr = &a;
    break;
  }
  if (!arr)
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] no single-component point data array in '%s'", filepath);
  const VtkType *vt = nullptr;
  {
    auto it = arr->attr.find("type");
    for (const auto &t : kVtkTypes)
      if (it != arr->attr.end() && it->second == t.name)
        vt = &t;
  }
  if (!vt)
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_VTI] unsupported vtk type '%s'",
        arr->attr.count("type") ? arr->attr.at("type").c_str() : "");
  if (vt->dvrType < 0)
    return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_VTI] vtk type %s maps to a 32/64-bit fixed-point type the field rejects",
        vt->name);
  const std::string format = arr->attr.count("format") ? arr->attr.at("format") : "ascii";
  const size_t want = (size_t)(nPoints * vt->size);
  std::vector<uint8_t> bytes;
  if (format == "ascii") {
    const size_t close = text.find("</DataArray", arr->end);
    if (arr->selfClosing || close == std::string::npos)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] ascii DataArray without content");
    bytes.resize(want);
    const char *p = text.c_str() + arr->end;
    const char *end = text.c_str() + close;
    for (uint64_t i = 0; i < nPoints; ++i) {
      char *next = nullptr;
      if (vt->kind == 'f') {
        const double v = std::strtod(p, &next);
        if (next == p || next > end)
          return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] ascii DataArray has fewer than %llu values", (unsigned long long)nPoints);
        if (vt->size == 4) {
          const float f = (float)v;
          std::memcpy(bytes.data() + i * 4, &f, 4);
        } else
          std::memcpy(bytes.data() + i * 8, &v, 8);
      } else {
        const long long v = std::strtoll(p, &next, 10);
        if (next == p || next > end)
          return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] ascii DataArray has fewer than %llu values", (unsigned long long)nPoints);
        if (vt->size == 1) {
          const uint8_t b = (uint8_t)v;
          bytes[i] = b;
        } else {
          const uint16_t h = (uint16_t)v;
          std::memcpy(bytes.data() + i * 2, &h, 2);
        }
      }
      p = next;
    }
  } else if (format == "binary") {
    const size_t close = text.find("</DataArray", arr->end);
    if (arr->selfClosing || close == std::string::npos)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] binary DataArray without content");
    const int rc = decodeBlock(text.c_str() + arr->end, close - arr->end, true, !compressor.empty(), hsz, want, bytes);
    if (rc != DVR_IMPORT_OK)
      return rc;
  } else if (format == "appended") {
    if (appendedData == std::string::npos)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] appended DataArray but no AppendedData section");
    const uint64_t offset = std::strtoull(arr->attr.count("offset") ? arr->attr.at("offset").c_str() : "0", nullptr, 10);
    if (appendedData + offset >= text.size())
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] appended offset beyond the end of the file");
    const int rc = decodeBlock(text.c_str() + appendedData + offset, text.size() - appendedData - offset, appendedBase64,
        !compressor.empty(), hsz, want, bytes);
    if (rc != DVR_IMPORT_OK)
      return rc;
  } else
    return fail(DVR_IMPORT_ERR_FORMAT, "[import_VTI] unknown DataArray format '%s'", format.c_str());
  if (bigEndian && format != "ascii")
    byteSwap(bytes.data(), nPoints, vt->size);

  o->kind = DVR_IMPORT_STRUCTURED;
  o->dataType = vt->dvrType;
  for (int i = 0; i < 3; ++i) {
    o->dims[i] = (uint32_t)dims[i];
    o->origin[i] = (float)origin[i]; // import_VTI.cpp:84-85
    o->spacing[i] = (float)spacing[i];
    o->headerSpacing[i] = spacing[i];
  }
  setName(o, fileOf(filepath));
  o->bytes = want;
  o->data = std::malloc(want ? want : 1);
  if (!o->data)
    return fail(DVR_IMPORT_ERR_IO, "[import_VTI] out of memory");
  std::memcpy(o->data, bytes.data(), want);
  return finishStructured(o);
}

// ---- NVDB (import_NVDB.cpp:17-115 through nanovdb::io::readGrid, io/IO.h:480-558, GridHandle.h:382-402) -------
constexpr uint64_t kMagicNumb = 0x304244566f6e614eull; // "NanoVDB0"
constexpr uint64_t kMagicGrid = 0x314244566f6e614eull; // "NanoVDB1"
constexpr uint64_t kMagicFile = 0x324244566f6e614eull; // "NanoVDB2"

template <typename T>
T rd(const uint8_t *p)
{
  T v;
  std::memcpy(&v, p, sizeof(T));
  return v;
}
template <typename T>
void wr(uint8_t *p, T v)
{
  std::memcpy(p, &v, sizeof(T));
}

bool versionCompatible(uint32_t v) { return (v >> 21) == 32u; } // Version::isCompatible, NanoVDB.h:709

// GridData::isValid (NanoVDB.h:1864-1875) on the first 672 bytes
bool gridDataValid(const uint8_t *g)
{
  const uint64_t magic = rd<uint64_t>(g);
  if (magic == kMagicGrid || rd<uint64_t>(g + 664) == kMagicGrid)
    return true;
  if (magic != kMagicNumb)
    return false;
  if (!versionCompatible(rd<uint32_t>(g + 16)))
    return false;
  const uint32_t idx = rd<uint32_t>(g + 24), cnt = rd<uint32_t>(g + 28);
  if (!(cnt > 0u && idx < cnt))
    return false;
  return rd<uint32_t>(g + 632) < 10u && rd<uint32_t>(g + 636) < 28u;
}

// tools::updateGridCount (GridChecksum.h:412-420)
void updateGridCount(uint8_t *g, uint32_t index, uint32_t count)
{
  if (rd<uint32_t>(g + 24) == index && rd<uint32_t>(g + 28) == count)
    return;
  wr<uint32_t>(g + 24, index);
  wr<uint32_t>(g + 28, count);
  const uint64_t checksum = rd<uint64_t>(g + 8);
  const uint32_t ver = rd<uint32_t>(g + 16);
  const bool newer = ver > ((32u << 21) | (6u << 10)); // crc32Head covers GridData+TreeData after 32.6.0
  if (checksum != ~0ull && newer) {
    const uint32_t head = (uint32_t)crc32(0L, g + 16, 672 + 64 - 16);
    wr<uint32_t>(g + 8, head);
  }
}

struct NvdbTree
{
  uint8_t *blob;
  uint8_t *root;
  uint32_t gridType;
  uint32_t tiles;
};

float leafValueAt(const NvdbTree &t, const uint8_t *leaf, uint32_t n)
{
  if (t.gridType == 1u)
    return rd<float>(leaf + 96 + 4 * (size_t)n);
  int b;
  switch (t.gridType) {
  case 13u: b = 2; break;
  case 14u: b = 3; break;
  case 15u: b = 4; break;
  default: b = leaf[15] >> 5; break;
  }
  uint32_t code = rd<uint32_t>(leaf + 96 + 4 * (size_t)(n >> (5 - b)));
  code >>= (n & ((32u >> b) - 1u)) << b;
  code &= (1u << (1u << b)) - 1u;
  return (float)code * rd<float>(leaf + 84) + rd<float>(leaf + 80);
}

// min / max over the ACTIVE values of the tree (what updateGridStats(StatsMode::MinMax) leaves in the root)
void activeMinMax(const NvdbTree &t, float &lo, float &hi)
{
  lo = FLT_MAX;
  hi = -FLT_MAX;
  auto on = [](const uint8_t *mask, uint32_t n) { return (rd<uint64_t>(mask + 8 * (n >> 6)) >> (n & 63)) & 1ull; };
  auto take = [&](float v) {
    lo = std::min(lo, v);
    hi = std::max(hi, v);
  };
  for (uint32_t ti = 0; ti < t.tiles; ++ti) {
    const uint8_t *tile = t.root + 64 + 32 * (size_t)ti;
    const int64_t child = rd<int64_t>(tile + 8);
    if (child == 0) {
      if (rd<uint32_t>(tile + 16)) // Tile::state
        take(rd<float>(tile + 20));
      continue;
    }
    const uint8_t *upper = t.root + child;
    for (uint32_t n = 0; n < 32768u; ++n) {
      const uint8_t *entry = upper + 8256 + 8 * (size_t)n;
      if (!on(upper + 32 + 4096, n)) {
        if (on(upper + 32, n))
          take(rd<float>(entry));
        continue;
      }
      const uint8_t *lower = upper + rd<int64_t>(entry);
      for (uint32_t m = 0; m < 4096u; ++m) {
        const uint8_t *e2 = lower + 1088 + 8 * (size_t)m;
        if (!on(lower + 32 + 512, m)) {
          if (on(lower + 32, m))
            take(rd<float>(e2));
          continue;
        }
        const uint8_t *leaf = lower + rd<int64_t>(e2);
        for (uint32_t w = 0; w < 8; ++w) {
          uint64_t bits = rd<uint64_t>(leaf + 16 + 8 * w);
          while (bits) {
            const uint32_t b = (uint32_t)__builtin_ctzll(bits);
            bits &= bits - 1;
            take(leafValueAt(t, leaf, w * 64 + b));
          }
        }
      }
    }
  }
}

int importNvdb(const char *filepath, DvrVolumeFile *o)
{
  const std::string file = fileOf(filepath);
  if (file.empty())
    return fail(DVR_IMPORT_ERR_ARGUMENT, "[import_NVDB] no file name in path '%s'", filepath);
  std::ifstream is(filepath, std::ios::in | std::ios::binary);
  if (!is.is_open())
    return fail(DVR_IMPORT_ERR_IO, "[import_NVDB] failed: Unable to open file named \"%s\" for input", filepath);
  is.seekg(0, std::ios::end);
  const uint64_t fileBytes = (uint64_t)is.tellg();
  is.seekg(0);
  std::vector<uint8_t> head(672, 0);
  is.read((char *)head.data(), (std::streamsize)std::min<uint64_t>(672, fileBytes));
  uint8_t *grid = nullptr;
  uint64_t gridSize = 0;
  if (fileBytes >= 672 && gridDataValid(head.data())) {
    // a raw grid buffer (GridHandle::read): grid #0 of the buffer
    if (rd<uint32_t>(head.data() + 28) == 0)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: stream does not contain a #0 grid");
    uint64_t at = 0;
    while (rd<uint32_t>(head.data() + 24) != 0u) { // skip to the grid whose index is 0
      at += rd<uint64_t>(head.data() + 32);
      if (at + 672 > fileBytes)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: raw buffer without grid #0");
      is.seekg((std::streamoff)at);
      is.read((char *)head.data(), 672);
    }
    gridSize = rd<uint64_t>(head.data() + 32);
    if (gridSize < 672 + 64 || at + gridSize > fileBytes)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: raw grid size exceeds the file");
    grid = (uint8_t *)std::malloc(gridSize);
    if (!grid)
      return fail(DVR_IMPORT_ERR_IO, "[import_NVDB] out of memory");
    is.seekg((std::streamoff)at);
    is.read((char *)grid, (std::streamsize)gridSize);
  } else {
    // segment file: FileHeader (16 B), per grid FileMetaData (176 B) + name, then the grids (IO.h:388-433)
    if (fileBytes < 16)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: Expected a NanoVDB file, but read a file of unknown type!");
    const uint64_t magic = rd<uint64_t>(head.data());
    if (magic != kMagicNumb && magic != kMagicFile) {
      if (magic == __builtin_bswap64(kMagicNumb) || magic == __builtin_bswap64(kMagicFile))
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: This nvdb file has reversed endianness");
      if (magic == kMagicGrid)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: Expected a NanoVDB file, but read a raw NanoVDB grid!");
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: Expected a NanoVDB file, but read a file of unknown type!");
    }
    if (!versionCompatible(rd<uint32_t>(head.data() + 8)))
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: Incompatible file format (NanoVDB major version != 32)");
    const uint16_t gridCount = rd<uint16_t>(head.data() + 12), codec = rd<uint16_t>(head.data() + 14);
    if (gridCount == 0)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: stream does not contain a #0 grid");
    uint64_t at = 16;
    uint64_t size0 = 0;
    for (uint16_t gi = 0; gi < gridCount; ++gi) {
      uint8_t meta[176];
      is.seekg((std::streamoff)at);
      is.read((char *)meta, sizeof(meta));
      if (!is)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: truncated grid meta data");
      if (gi == 0)
        size0 = rd<uint64_t>(meta);
      at += 176 + rd<uint32_t>(meta + 136); // + nameSize
    }
    gridSize = size0;
    if (gridSize < 672 + 64 || gridSize > (1ull << 40) || at > fileBytes)
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: implausible grid size");
    // The size of the allocation comes from the file: before trusting it, hold it against what the file can deliver —
    // the bytes that are left (Codec::NONE) or what a zlib stream of the stored length can expand to (deflate's limit
    // is 1032:1; Codec::ZIP).
    uint64_t csz = 0;
    is.seekg((std::streamoff)at);
    if (codec == 0) {
      if (gridSize > fileBytes - at)
        return fail(DVR_IMPORT_ERR_IO, "[import_NVDB] failed: Failed to read Tree from file");
    } else if (codec == 1) {
      is.read((char *)&csz, 8);
      if (!is || at + 8 > fileBytes || csz > fileBytes - at - 8)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: truncated ZIP stream");
      if (gridSize / 1032u > csz + 64u)
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: UNZIP failed on byte size");
    } else
      return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_NVDB] failed: BLOSC compression codec was disabled during build");
    grid = (uint8_t *)std::malloc(gridSize);
    if (!grid)
      return fail(DVR_IMPORT_ERR_IO, "[import_NVDB] out of memory");
    if (codec == 0) { // Codec::NONE
      is.read((char *)grid, (std::streamsize)gridSize);
      if (!is) {
        std::free(grid);
        return fail(DVR_IMPORT_ERR_IO, "[import_NVDB] failed: Failed to read Tree from file");
      }
    } else if (codec == 1) { // Codec::ZIP: u64 compressed size + one zlib stream (IO.h:314-327)
      std::vector<uint8_t> tmp(csz);
      is.read((char *)tmp.data(), (std::streamsize)csz);
      uLongf n = (uLongf)gridSize;
      const int st = uncompress(grid, &n, tmp.data(), (uLong)csz);
      if (!is || st != Z_OK || (uint64_t)n != gridSize) {
        std::free(grid);
        return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: UNZIP failed on byte size");
      }
    } else {
      std::free(grid);
      return fail(DVR_IMPORT_ERR_UNSUPPORTED, "[import_NVDB] failed: BLOSC compression codec was disabled during build");
    }
    if (!gridDataValid(grid) && rd<uint64_t>(grid) != kMagicNumb) {
      std::free(grid);
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: the segment does not hold a NanoVDB grid");
    }
  }
  updateGridCount(grid, 0u, 1u);

  // min/max: the root's, after updateGridStats(MinMax) when the grid carries none (import_NVDB.cpp:30-88)
  NvdbTree t;
  t.blob = grid;
  t.gridType = rd<uint32_t>(grid + 636);
  const int64_t rootOff = 672 + rd<int64_t>(grid + 672 + 24);
  const bool supported = t.gridType == 1u || (t.gridType >= 13u && t.gridType <= 16u);
  // every offset the min/max pass below (and later the device) follows comes from the file: check them all first
  if (supported) {
    const int bad = dvr_nvdb_validate_tree(grid, gridSize);
    if (bad != 0) {
      std::free(grid);
      return fail(DVR_IMPORT_ERR_FORMAT, "[import_NVDB] failed: corrupt NanoVDB tree (node offsets leave the grid buffer, code %d)", bad);
    }
  }
  o->hasValueRange = 0;
  if (supported && rootOff >= 736 && (uint64_t)rootOff + 64 <= gridSize) {
    t.root = grid + rootOff;
    t.tiles = rd<uint32_t>(t.root + 24);
    const uint32_t flags = rd<uint32_t>(grid + 20);
    if (!(flags & (1u << 2))) { // GridFlags::HasMinMax
      float lo, hi;
      activeMinMax(t, lo, hi);
      if (lo <= hi) {
        wr<float>(t.root + 32, lo);
        wr<float>(t.root + 36, hi);
        wr<uint32_t>(grid + 20, flags | (1u << 2));
      }
    }
    const float lo = rd<float>(t.root + 32), hi = rd<float>(t.root + 36);
    if (lo <= hi) {
      o->valueRange[0] = lo;
      o->valueRange[1] = hi;
      o->hasValueRange = 1;
    }
  }
  o->kind = DVR_IMPORT_NANOVDB;
  o->dataType = -1;
  const uint8_t *rootBBox = grid + rootOff;
  if (supported && rootOff >= 736 && (uint64_t)rootOff + 64 <= gridSize)
    for (int i = 0; i < 3; ++i) {
      const int32_t a = rd<int32_t>(rootBBox + 4 * i), b = rd<int32_t>(rootBBox + 12 + 4 * i);
      o->dims[i] = b >= a ? (uint32_t)(b - a + 1) : 0u;
    }
  for (int i = 0; i < 3; ++i) {
    o->origin[i] = (float)rd<double>(grid + 560 + 8 * i); // world bounding box minimum
    o->spacing[i] = (float)rd<double>(grid + 608 + 8 * i);
    o->headerSpacing[i] = rd<double>(grid + 608 + 8 * i);
  }
  setName(o, file);
  o->data = grid;
  o->bytes = gridSize;
  return DVR_IMPORT_OK;
}

} // namespace

extern "C" {

const char *dvr_import_last_error(void) { return g_err; }

#define IMPORT_ENTRY(fn, impl)                                                                                    \
  int fn(const char *path, DvrVolumeFile *out)                                                                      \
  {                                                                                                                 \
    if (!path || !out)                                                                                              \
      return fail(DVR_IMPORT_ERR_ARGUMENT, #fn ": null argument");                                                  \
    initOut(out);                                                                                                   \
    g_err[0] = 0;                                                                                                   \
    const int rc = impl(path, out);                                                                                 \
    if (rc != DVR_IMPORT_OK) {                                                                                      \
      if (out->data)                                                                                                \
        std::free(out->data);                                                                                       \
      initOut(out);                                                                                                 \
    }                                                                                                               \
    return rc;                                                                                                      \
  }

IMPORT_ENTRY(dvr_import_raw, importRaw)
IMPORT_ENTRY(dvr_import_mhd, importMhd)
IMPORT_ENTRY(dvr_import_vti, importVti)
IMPORT_ENTRY(dvr_import_nvdb, importNvdb)

// import_volume.cpp:21-37 (".flash" and ".vtu" are AMR / unstructured fields: outside the DVR path of this device)
int dvr_import_volume(const char *path, DvrVolumeFile *out)
{
  if (!path || !out)
    return fail(DVR_IMPORT_ERR_ARGUMENT, "dvr_import_volume: null argument");
  const std::string ext = extensionOf(path);
  if (ext == ".raw")
    return dvr_import_raw(path, out);
  if (ext == ".nvdb")
    return dvr_import_nvdb(path, out);
  if (ext == ".mhd")
    return dvr_import_mhd(path, out);
  if (ext == ".vti")
    return dvr_import_vti(path, out);
  initOut(out);
  return fail(DVR_IMPORT_ERR_ARGUMENT, "[import_volume] no loader for file type '%s'", ext.c_str());
}

void dvr_import_free(DvrVolumeFile *f)
{
  if (f && f->data) {
    std::free(f->data);
    f->data = nullptr;
    f->bytes = 0;
  }
}

int dvr_compute_scalar_range(const void *hostData, int32_t dataType, uint64_t n, float out[2])
{
  if (!out)
    return fail(DVR_IMPORT_ERR_ARGUMENT, "dvr_compute_scalar_range: null argument");
  return scalarRange(hostData, dataType, n, out);
}

} // extern "C"
