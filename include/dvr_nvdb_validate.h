/*
 * dvr_nvdb_validate.h — bounds validation of a serialized NanoVDB grid before anything walks it.
 *
 * A NanoVDB grid is position-independent: every node is found through byte offsets stored in the buffer itself
 * (TreeData::mNodeOffset, root tile child offsets, internal-node child tables).  A truncated or crafted file can
 * point those anywhere.  This header-only, host-only walk checks that every node a reader can reach — root tile
 * table, upper 32^3 nodes, lower 16^3 nodes, leaves (float or Fp4/Fp8/Fp16/FpN) — lies inside [0, gridSize), so that
 * the host min/max pass of the importer (visrtx_b200/importers/volume_import.cpp) and the device tree walk / brick
 * gather (visrtx_b200/csrc/dvr_nanovdb.cuh) never dereference outside the buffer.  Layout constants are those of
 * dvr_nanovdb.cuh (NanoVDB 32.x: GridData 672 B, TreeData 64 B, RootData<float> 64 B + 32-byte tiles).
 *
 * The reference trusts the NanoVDB library's own GridHandle / validation on this path
 * (devices/rtx/scene/volume/spatial_field/NvdbRegularField.cpp:64-113).
 */
#ifndef DVR_NVDB_VALIDATE_H
#define DVR_NVDB_VALIDATE_H

#include <stdint.h>
#include <string.h>

static inline int dvr_nvdb_in_range(uint64_t gridSize, int64_t off, uint64_t bytes)
{
  return off >= 0 && (uint64_t)off <= gridSize && bytes <= gridSize - (uint64_t)off;
}

/* Returns 0 when every reachable node lies inside the buffer, else a negative code:
 *  -1 header / root out of range, -2 tile table, -3 upper node, -4 lower node, -5 leaf, -6 unsupported grid type.
 * `grid` points at `gridSize` readable bytes on the HOST. */
static inline int dvr_nvdb_validate_tree(const uint8_t *grid, uint64_t gridSize)
{
  if (!grid || gridSize < 672 + 64 + 64)
    return -1;
  uint32_t gridType;
  memcpy(&gridType, grid + 636, 4);
  const int isFloat = gridType == 1u;
  if (!isFloat && !(gridType >= 13u && gridType <= 16u))
    return -6;
  int64_t rootRel;
  memcpy(&rootRel, grid + 672 + 24, 8);
  const int64_t rootOff = 672 + rootRel;
  if (rootOff < 736 || !dvr_nvdb_in_range(gridSize, rootOff, 64))
    return -1;
  uint32_t tiles;
  memcpy(&tiles, grid + rootOff + 24, 4);
  if (!dvr_nvdb_in_range(gridSize, rootOff + 64, (uint64_t)tiles * 32u))
    return -2;
  const uint64_t upperBytes = 8256u + 8u * 32768u, lowerBytes = 1088u + 8u * 4096u;
  for (uint32_t ti = 0; ti < tiles; ++ti) {
    int64_t child;
    memcpy(&child, grid + rootOff + 64 + 32 * (uint64_t)ti + 8, 8);
    if (child == 0)
      continue;
    const int64_t upperOff = rootOff + child;
    if (!dvr_nvdb_in_range(gridSize, upperOff, upperBytes))
      return -3;
    const uint8_t *upper = grid + upperOff;
    for (uint32_t n = 0; n < 32768u; ++n) {
      uint64_t w;
      memcpy(&w, upper + 32 + 4096 + 8 * (n >> 6), 8);
      if (!((w >> (n & 63u)) & 1ull))
        continue;
      int64_t rel;
      memcpy(&rel, upper + 8256 + 8 * (uint64_t)n, 8);
      const int64_t lowerOff = upperOff + rel;
      if (!dvr_nvdb_in_range(gridSize, lowerOff, lowerBytes))
        return -4;
      const uint8_t *lower = grid + lowerOff;
      for (uint32_t m = 0; m < 4096u; ++m) {
        uint64_t w2;
        memcpy(&w2, lower + 32 + 512 + 8 * (m >> 6), 8);
        if (!((w2 >> (m & 63u)) & 1ull))
          continue;
        int64_t rel2;
        memcpy(&rel2, lower + 1088 + 8 * (uint64_t)m, 8);
        const int64_t leafOff = lowerOff + rel2;
        if (!dvr_nvdb_in_range(gridSize, leafOff, 96))
          return -5;
        uint64_t leafBytes = 96u + 512u * 4u; /* float leaf */
        if (!isFloat) {
          uint32_t log2Bits = gridType == 13u ? 2u : gridType == 14u ? 3u : gridType == 15u ? 4u
                                                                                             : (uint32_t)(grid[leafOff + 15] >> 5);
          if (log2Bits > 4u)
            return -5;
          leafBytes = 96u + (64u << log2Bits); /* 512 codes of 2^log2Bits bits */
        }
        if (!dvr_nvdb_in_range(gridSize, leafOff, leafBytes))
          return -5;
      }
    }
  }
  return 0;
}

#endif /* DVR_NVDB_VALIDATE_H */
