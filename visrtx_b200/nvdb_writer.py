"""Minimal NanoVDB writer: serialises a dense float32 block as a sparse NanoVDB grid — ``GridType::Float`` or
one of the quantised types Fp4 / Fp8 / Fp16 / FpN (per-leaf ``min + code * quantum`` codes of 4 / 8 / 16 / a
per-leaf 1..16 bits).

Used to create synthetic "nanovdb" spatial fields (BASELINE config C5: a fog sphere) without the NanoVDB
library: the layout follows the published NanoVDB 32.7 binary format — GridData (672 B), TreeData (64 B),
RootData<float> (64 B) + 32-byte tiles with 64-bit keys, upper 32^3 nodes (270 400 B), lower 16^3 nodes
(33 856 B), 8^3 float leaves (2 144 B), breadth-first.  Only what a reader needs is filled in (child masks,
child offsets, leaf values and value masks, bounding boxes, the index->world map); statistics are min/max
only and the checksum is disabled, as NanoVDB itself allows.

The output is checked against the reference's vendored NanoVDB in tests/test_nanovdb_host.py: the real
``grid->isValid()``, ``worldBBox``, ``activeVoxelCount`` and ``SampleFromVoxels`` read it back.
"""
from __future__ import annotations

import numpy as np

_GRID, _TREE, _ROOT, _TILE = 672, 64, 64, 32
_UPPER, _LOWER, _LEAF = 270400, 33856, 2144
_UPPER_TABLE, _LOWER_TABLE, _LEAF_VALUES = 8256, 1088, 96
_MAGIC_NUMB = 0x304244566F6E614E  # "NanoVDB0"
_MAGIC_GRID = 0x314244566F6E614E  # "NanoVDB1"
GRID_TYPES = {"float": 1, "fp4": 13, "fp8": 14, "fp16": 15, "fpn": 16}  # nanovdb::GridType
_FIXED_BITS = {"fp4": 4, "fp8": 8, "fp16": 16}


def _quantise_leaves(leaf_vals: np.ndarray, codec: str, tolerance: float):
    """Per leaf: (bits, minimum, quantum, codes) with value ~= minimum + code * quantum (LeafFnBase, NanoVDB.h).
    Fixed types use their bit width; FpN picks per leaf the smallest of 1/2/4/8/16 bits whose reconstruction
    error stays within ``tolerance`` (NanoVDB's AbsDiff oracle does the same job)."""
    n = len(leaf_vals)
    vmin = leaf_vals.min(axis=1).astype(np.float32)
    vmax = leaf_vals.max(axis=1).astype(np.float32)
    span = (vmax - vmin).astype(np.float32)

    def encode(bits):
        units = np.float32((1 << bits) - 1)
        quantum = (span / units).astype(np.float32)
        enc = np.where(span > 0, units / np.where(span > 0, span, 1), 0).astype(np.float32)
        codes = np.floor(enc[:, None] * (leaf_vals - vmin[:, None]) + np.float32(0.5))
        codes = np.clip(codes, 0, float(units)).astype(np.uint32)
        return quantum, codes

    if codec in _FIXED_BITS:
        b = _FIXED_BITS[codec]
        q, c = encode(b)
        return np.full(n, b, np.int32), vmin, q, c
    bits = np.full(n, 16, np.int32)
    quantum = np.zeros(n, np.float32)
    codes = np.zeros((n, 512), np.uint32)
    done = np.zeros(n, bool)
    for b in (1, 2, 4, 8, 16):
        q, c = encode(b)
        err = np.abs(vmin[:, None] + c.astype(np.float32) * q[:, None] - leaf_vals).max(axis=1)
        take = ~done & ((err <= tolerance) | (b == 16))
        bits[take], quantum[take], codes[take] = b, q[take], c[take]
        done |= take
    return bits, vmin, quantum, codes


def _pack_codes(codes: np.ndarray, bits: int) -> np.ndarray:
    """(m,512) codes -> (m, 64*bits) bytes, code i in bits [i*bits, (i+1)*bits) of the little-endian stream."""
    if bits == 16:
        return np.ascontiguousarray(codes.astype("<u2")).view(np.uint8).reshape(len(codes), 1024)
    if bits == 8:
        return codes.astype(np.uint8)
    per = 8 // bits
    c = codes.reshape(len(codes), 512 // per, per).astype(np.uint8)
    out = np.zeros(c.shape[:2], np.uint8)
    for k in range(per):
        out |= c[:, :, k] << (k * bits)
    return out


def write_float_grid(values: np.ndarray, index_origin=(0, 0, 0), voxel_size: float = 1.0, world_origin=(0.0, 0.0, 0.0),
                     background: float = 0.0, name: str = "density", grid_class: int = 2, codec: str = "float",
                     tolerance: float = 1e-3) -> np.ndarray:
    """values[x, y, z] (float32) holds the voxel at index ``index_origin + (x, y, z)``; voxels equal to
    ``background`` are inactive and are not stored.  world = voxel_size * index + world_origin.
    Returns the serialised grid as a uint8 array (32-byte multiple)."""
    v = np.ascontiguousarray(values, dtype=np.float32)
    if v.ndim != 3:
        raise ValueError("values must be a 3-D array indexed [x, y, z]")
    o = np.asarray(index_origin, dtype=np.int64)
    # pad to leaf (8^3) alignment in GLOBAL index space
    lo = (o // 8) * 8
    hi = -((-(o + np.asarray(v.shape))) // 8) * 8
    padded = np.full(tuple((hi - lo).tolist()), background, dtype=np.float32)
    s = o - lo
    padded[s[0]:s[0] + v.shape[0], s[1]:s[1] + v.shape[1], s[2]:s[2] + v.shape[2]] = v
    nb = (hi - lo) // 8
    blocks = padded.reshape(nb[0], 8, nb[1], 8, nb[2], 8).transpose(0, 2, 4, 1, 3, 5)  # [bx,by,bz,8,8,8]
    active = blocks != np.float32(background)
    has = active.reshape(nb[0], nb[1], nb[2], -1).any(axis=-1)
    bidx = np.argwhere(has)  # leaf block indices, sorted x-major (NanoVDB's breadth-first order within a parent)
    if len(bidx) == 0:
        raise ValueError("the grid has no active voxel")
    leaf_origin = bidx * 8 + lo  # global index coords of each leaf

    # parents
    lower_key = leaf_origin >> 7
    upper_key = leaf_origin >> 12
    lowers, lower_of_leaf = np.unique(lower_key, axis=0, return_inverse=True)
    uppers, upper_of_lower = np.unique(lowers >> 5, axis=0, return_inverse=True)
    lower_of_leaf = lower_of_leaf.ravel()
    upper_of_lower = upper_of_lower.ravel()
    n_leaf, n_lower, n_upper = len(leaf_origin), len(lowers), len(uppers)

    if codec not in GRID_TYPES:
        raise ValueError(f"codec must be one of {sorted(GRID_TYPES)}")
    leaf_vals = blocks[bidx[:, 0], bidx[:, 1], bidx[:, 2]].reshape(n_leaf, 512)
    if codec == "float":
        leaf_size = np.full(n_leaf, _LEAF, np.int64)
    else:
        q_bits, q_min, q_quantum, q_codes = _quantise_leaves(leaf_vals, codec, tolerance)
        leaf_size = 96 + 64 * q_bits.astype(np.int64)
    leaf_off = np.concatenate([[0], np.cumsum(leaf_size)[:-1]]).astype(np.int64)  # relative to the first leaf

    off_root = _GRID + _TREE
    off_upper = off_root + _ROOT + _TILE * n_upper
    off_lower = off_upper + _UPPER * n_upper
    off_leaf = off_lower + _LOWER * n_lower
    total = off_leaf + int(leaf_size.sum())
    buf = np.zeros(total, dtype=np.uint8)

    def put(off, arr):
        b = np.ascontiguousarray(arr).view(np.uint8).ravel()
        buf[off:off + b.size] = b

    act_vals = v[v != np.float32(background)]
    vmin, vmax = float(act_vals.min()), float(act_vals.max())
    ai = np.argwhere(v != np.float32(background))
    bb_min = ai.min(axis=0) + o
    bb_max = ai.max(axis=0) + o

    # ---- GridData
    put(0, np.array([_MAGIC_NUMB], np.uint64))
    put(8, np.array([0xFFFFFFFFFFFFFFFF], np.uint64))  # checksum disabled
    put(16, np.array([(32 << 21) | (7 << 10) | 0], np.uint32))  # version 32.7.0
    put(20, np.array([0], np.uint32))  # flags: none claimed (no bbox/minmax guarantees)
    put(24, np.array([0, 1], np.uint32))  # grid index, grid count
    put(32, np.array([total], np.uint64))
    nm = name.encode()[:255]
    buf[40:40 + len(nm)] = np.frombuffer(nm, np.uint8)
    s_ = float(voxel_size)
    matf = np.diag([s_, s_, s_]).astype(np.float32).ravel()
    invf = np.diag([1.0 / np.float32(s_)] * 3).astype(np.float32).ravel()
    m = 296
    put(m, matf)
    put(m + 36, invf)
    put(m + 72, np.asarray(world_origin, np.float32))
    put(m + 84, np.array([1.0], np.float32))
    put(m + 88, np.diag([s_, s_, s_]).astype(np.float64).ravel())
    put(m + 160, np.diag([1.0 / s_] * 3).astype(np.float64).ravel())
    put(m + 232, np.asarray(world_origin, np.float64))
    put(m + 256, np.array([1.0], np.float64))
    wmin = bb_min.astype(np.float64) * s_ + np.asarray(world_origin, np.float64)
    wmax = (bb_max.astype(np.float64) + 1.0) * s_ + np.asarray(world_origin, np.float64)
    put(560, np.concatenate([wmin, wmax]))
    put(608, np.array([s_, s_, s_], np.float64))
    put(632, np.array([grid_class, GRID_TYPES[codec]], np.uint32))  # GridClass (2 = FogVolume), GridType
    put(640, np.array([total], np.int64))  # blind metadata offset = grid size (none)
    put(648, np.array([0, 0], np.uint32))
    put(656, np.array([0, _MAGIC_GRID], np.uint64))

    # ---- TreeData (offsets relative to the tree)
    put(_GRID, np.array([off_leaf - _GRID, off_lower - _GRID, off_upper - _GRID, off_root - _GRID], np.int64))
    put(_GRID + 32, np.array([n_leaf, n_lower, n_upper, 0, 0, 0], np.uint32))
    put(_GRID + 56, np.array([int(active.sum())], np.uint64))

    # ---- RootData + tiles
    put(off_root, np.concatenate([bb_min, bb_max]).astype(np.int32))
    put(off_root + 24, np.array([n_upper], np.uint32))
    put(off_root + 28, np.array([background, vmin, vmax, 0.0, 0.0], np.float32))
    for i, uk in enumerate(uppers):
        t = off_root + _ROOT + _TILE * i
        ux, uy, uz = (int(c) & 0xFFFFF for c in uk)  # uint32(coord) >> 12 keeps 20 bits
        key = uz | (uy << 21) | (ux << 42)
        put(t, np.array([key], np.uint64))
        put(t + 8, np.array([off_upper + _UPPER * i - off_root], np.int64))  # child offset from the root

    def set_bits(mask_off, idx, n_words):
        words = np.zeros(n_words, np.uint64)
        np.bitwise_or.at(words, (idx >> 6).astype(np.int64), np.uint64(1) << (idx & 63).astype(np.uint64))
        put(mask_off, words)

    # ---- upper nodes
    bgbits = np.array([background], np.float32).view(np.uint32)[0]
    for i, uk in enumerate(uppers):
        base = off_upper + _UPPER * i
        org = uk.astype(np.int64) * 4096
        put(base, np.concatenate([org, org + 4095]).astype(np.int32))
        kids = np.where(upper_of_lower == i)[0]
        lk = lowers[kids]
        n = (((lk[:, 0] & 31) << 10) | ((lk[:, 1] & 31) << 5) | (lk[:, 2] & 31)).astype(np.int64)
        set_bits(base + 32 + 4096, n, 512)
        table = np.zeros(32768, np.int64)
        table[:] = int(bgbits)  # tile value in the low 4 bytes
        table[n] = (off_lower + _LOWER * kids) - base
        put(base + _UPPER_TABLE, table)
        put(base + 32 + 8192, np.array([vmin, vmax, 0.0, 0.0], np.float32))

    # ---- lower nodes
    for i, lk in enumerate(lowers):
        base = off_lower + _LOWER * i
        org = lk.astype(np.int64) * 128
        put(base, np.concatenate([org, org + 127]).astype(np.int32))
        kids = np.where(lower_of_leaf == i)[0]
        lo3 = leaf_origin[kids] >> 3
        n = (((lo3[:, 0] & 15) << 8) | ((lo3[:, 1] & 15) << 4) | (lo3[:, 2] & 15)).astype(np.int64)
        set_bits(base + 32 + 512, n, 64)
        table = np.zeros(4096, np.int64)
        table[:] = int(bgbits)
        table[n] = (off_leaf + leaf_off[kids]) - base
        put(base + _LOWER_TABLE, table)
        put(base + 32 + 1024, np.array([vmin, vmax, 0.0, 0.0], np.float32))

    # ---- leaves (vectorised per leaf size)
    leaf_act = active[bidx[:, 0], bidx[:, 1], bidx[:, 2]].reshape(n_leaf, 512)
    bits = np.packbits(leaf_act.astype(np.uint8), axis=1, bitorder="little")  # bit n of word n>>6 == voxel n
    head = np.zeros((n_leaf, 96), np.uint8)
    head[:, 0:12] = np.ascontiguousarray(leaf_origin.astype(np.int32)).view(np.uint8).reshape(n_leaf, 12)
    head[:, 12:15] = 7
    head[:, 16:80] = bits
    if codec == "float":
        mm = np.stack([leaf_vals.min(axis=1), leaf_vals.max(axis=1), np.zeros(n_leaf, np.float32),
                       np.zeros(n_leaf, np.float32)], axis=1).astype(np.float32)
        head[:, 80:96] = np.ascontiguousarray(mm).view(np.uint8).reshape(n_leaf, 16)
        leaves = buf[off_leaf:].reshape(n_leaf, _LEAF)
        leaves[:, :96] = head
        leaves[:, _LEAF_VALUES:] = np.ascontiguousarray(leaf_vals).view(np.uint8).reshape(n_leaf, 2048)
        return buf
    # quantised leaf = LeafFnBase: ... float mMinimum @80, float mQuantum @84, u16 mMin,mMax,mAvg,mDev @88
    log2b = {1: 0, 2: 1, 4: 2, 8: 3, 16: 4}
    head[:, 15] = np.array([log2b[int(b)] << 5 for b in q_bits], np.uint8) if codec == "fpn" else 0
    head[:, 80:84] = np.ascontiguousarray(q_min).view(np.uint8).reshape(n_leaf, 4)
    head[:, 84:88] = np.ascontiguousarray(q_quantum).view(np.uint8).reshape(n_leaf, 4)
    cmax = ((1 << q_bits.astype(np.int64)) - 1).astype(np.uint16)
    head[:, 90:92] = np.ascontiguousarray(cmax).view(np.uint8).reshape(n_leaf, 2)  # mMax (mMin = 0)
    for b in np.unique(q_bits):
        sel = np.where(q_bits == b)[0]
        packed = _pack_codes(q_codes[sel], int(b))
        rows = (off_leaf + leaf_off[sel])[:, None] + np.arange(96 + 64 * int(b))[None, :]
        buf[rows] = np.concatenate([head[sel], packed], axis=1)
    return buf


def fog_sphere(radius: float = 100.0, voxel_size: float = 1.0, half_width: float = 3.0, codec: str = "float",
               tolerance: float = 1e-3) -> np.ndarray:
    """A fog-volume sphere like nanovdb::tools::createFogVolumeSphere: density 1 inside, falling linearly to 0
    over ``half_width`` voxels at the surface, 0 (background, inactive) outside."""
    r = radius / voxel_size
    n = int(np.ceil(r)) + 1
    c = np.arange(-n, n + 1, dtype=np.float32)
    x, y, z = np.meshgrid(c, c, c, indexing="ij", sparse=True)
    d = r - np.sqrt(x * x + y * y + z * z)  # signed distance to the surface in voxels (positive inside)
    dens = np.clip(d / np.float32(half_width), 0.0, 1.0).astype(np.float32)
    return write_float_grid(dens, index_origin=(-n, -n, -n), voxel_size=voxel_size, name="sphere_fog", codec=codec,
                            tolerance=tolerance)
