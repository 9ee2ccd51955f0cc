"""Generates tests/golden/import_fixtures.npz IN THE BUILD CONTAINER (needs oracle/_ref/libref_host.so):

    python tests/golden/make_import_fixtures.py

.nvdb files written by the reference's own NanoVDB I/O (nanovdb::io::writeGrid with codec NONE / ZIP, and bare
grid buffers), stored as byte arrays, together with what the reference's import path reads back from each of them
(nanovdb::io::readGrid -> grid bytes, root min/max after updateGridStats when needed; import_NVDB.cpp:25-88).
A grid without min/max flags (written by visrtx_b200.nvdb_writer) exercises the updateGridStats branch.
"""
import ctypes as C
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import oracle_binding as ob  # noqa: E402
from visrtx_b200 import nvdb_writer  # noqa: E402

CASES = [  # name, grid type, radius, codec (0 NONE, 1 ZIP), raw
    ("fog_float_none.nvdb", "float", 5.0, 0, 0),
    ("fog_float_zip.nvdb", "float", 7.0, 1, 0),
    ("fog_fp8_zip.nvdb", "fp8", 6.0, 1, 0),
    ("fog_fp4_none.nvdb", "fp4", 5.0, 0, 0),
    ("fog_fp16_raw.nvdb", "fp16", 5.0, 0, 1),
    ("fog_fpn_zip.nvdb", "fpn", 6.0, 1, 0),
]


def read_back(lib, path):
    size = C.c_size_t()
    mm = (C.c_float * 2)()
    assert lib.refhost_nvdb_read_file(path.encode(), None, C.c_size_t(0), C.byref(size), mm) == 0
    buf = np.zeros(size.value, np.uint8)
    assert lib.refhost_nvdb_read_file(path.encode(), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.byref(size),
                                      mm) == 0
    return buf, np.array([mm[0], mm[1]], np.float32)


def main():
    lib = ob.refhost()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, gt, radius, codec, raw in CASES:
            p = os.path.join(tmp, name)
            rc = lib.refhost_nvdb_write_file(p.encode(), C.c_uint(ob.NVDB_GRID_TYPES[gt]), C.c_double(radius),
                                             C.c_int(codec), C.c_int(raw))
            assert rc == 0, (name, rc)
            out["file/" + name] = np.fromfile(p, np.uint8)
            grid, mm = read_back(lib, p)
            if gt != "fpn":  # NanoVDB leaves FpN allocation slack uninitialised: only its samples are reproducible
                out["grid/" + name] = grid
            out["minmax/" + name] = mm
            print(name, out["file/" + name].nbytes, "bytes on disk ->", grid.nbytes, "grid bytes, min/max", mm)
        # two grids in one segment file (writeGrids): the importer takes grid #0
        for name, codec in (("two_grids_zip.nvdb", 1),):
            p = os.path.join(tmp, name)
            assert lib.refhost_nvdb_write_two_grid_file(p.encode(), C.c_double(5.0), C.c_int(codec)) == 0
            out["file/" + name] = np.fromfile(p, np.uint8)
            grid, mm = read_back(lib, p)
            out["grid/" + name] = grid
            out["minmax/" + name] = mm
            print(name, out["file/" + name].nbytes, "bytes on disk ->", grid.nbytes, "grid bytes (grid #0), min/max", mm)
        # a grid that carries no min/max statistics: raw buffer from our own writer (flags == 0)
        rng = np.random.default_rng(23)
        dense = (rng.random((9, 14, 11)) * 0.8 + 0.1).astype(np.float32)
        dense[rng.random(dense.shape) < 0.5] = 0.0
        blob = nvdb_writer.write_float_grid(dense, index_origin=(-3, 2, 5), voxel_size=0.5)
        name = "own_writer_nostats_raw.nvdb"
        p = os.path.join(tmp, name)
        blob.tofile(p)
        out["file/" + name] = blob
        _, mm = read_back(lib, p)
        out["minmax/" + name] = mm
        print(name, blob.nbytes, "bytes, min/max after updateGridStats", mm,
              "(active values:", dense[dense != 0].min(), dense.max(), ")")
    np.savez_compressed(os.path.join(HERE, "import_fixtures.npz"), **out)


if __name__ == "__main__":
    main()
