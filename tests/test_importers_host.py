"""CPU suite: the volume-file importers of include/dvr_import.h against
  * .nvdb files written by the reference's own NanoVDB I/O and what its import path reads back from them
    (tests/golden/import_fixtures.npz, generator tests/golden/make_import_fixtures.py; regenerated live when
    oracle/_ref/libref_host.so is present),
  * the file-name / header conventions of import_RAW.cpp and import_MHD.cpp, quirks included,
  * VTK XML ImageData files in every encoding the format defines (written here from the published format —
    VTK itself is a third-party dependency that is not in the reference tree: VTI parity is unpinned against VTK),
  * computeScalarRange's normalised extrema.
"""
import base64
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

import oracle_binding as ob
from visrtx_b200 import capi, importers as I, nvdb_writer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIX = np.load(os.path.join(GOLD, "import_fixtures.npz"))
NVDB_FILES = sorted(k[5:] for k in FIX.files if k.startswith("file/"))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "include", "dvr_import.h")).read()
    import re
    declared = set(re.findall(r"\b(dvr_[a-z_]+)\s*\(", hdr))
    assert declared == set(I.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(I.lib, name), name
    assert C.sizeof(I.DvrVolumeFile) == 4 + 4 + 12 + 12 + 12 + 4 + 24 + 8 + 4 + 4 + 8 + 8 + 256


# ------------------------------------------------------------------------------------------------- NVDB
@pytest.mark.parametrize("name", NVDB_FILES)
def test_import_nvdb_matches_the_reference_reader(name, tmp_path):
    p = tmp_path / name
    FIX["file/" + name].tofile(p)
    vf = I.import_nvdb(str(p))
    assert vf.kind == I.NANOVDB and vf.name == name
    if "grid/" + name in FIX.files:  # the grid buffer nanovdb::io::readGrid returns, byte for byte
        assert np.array_equal(vf.data, FIX["grid/" + name])
    mm = FIX["minmax/" + name]
    assert vf.has_value_range
    assert vf.value_range[0] == mm[0] and vf.value_range[1] == mm[1]
    # and the grid is one the field accepts: a single grid, index 0
    assert vf.data[24:32].view(np.uint32).tolist() == [0, 1]
    if ob.have_ref_host():
        assert ob.nvdb_is_valid_reference(np.ascontiguousarray(vf.data))


def test_import_nvdb_samples_like_the_source_grid(tmp_path):
    """End to end: file written by NanoVDB (ZIP) -> importer -> O-cpu sampler == sampling the in-memory grid."""
    if not ob.have_ref_host():
        pytest.skip("oracle/_ref/libref_host.so not built")
    name = "fog_fp8_zip.nvdb"
    p = tmp_path / name
    FIX["file/" + name].tofile(p)
    vf = I.import_nvdb(str(p))
    direct = ob.nvdb_fog_sphere_typed("fp8", 6.0)
    rng = np.random.default_rng(2)
    xyz = (rng.random((2000, 3)) * 16 - 8).astype(np.float32)
    assert np.array_equal(ob.nvdb_sample_oracle(np.ascontiguousarray(vf.data), xyz), ob.nvdb_sample_oracle(direct, xyz))


def test_import_nvdb_fixture_is_what_the_reference_writes(tmp_path):
    if not ob.have_ref_host():
        pytest.skip("oracle/_ref/libref_host.so not built")
    lib = ob.refhost()
    p = tmp_path / "live.nvdb"
    assert lib.refhost_nvdb_write_file(str(p).encode(), C.c_uint(1), C.c_double(5.0), C.c_int(0), C.c_int(0)) == 0
    assert np.array_equal(np.fromfile(p, np.uint8), FIX["file/fog_float_none.nvdb"])


def test_import_nvdb_multi_grid_raw_buffer_and_errors(tmp_path):
    a = nvdb_writer.fog_sphere(4.0)
    b = nvdb_writer.fog_sphere(5.0, codec="fp8")
    for i, g in enumerate((a, b)):  # two grids back to back: indices 0,1 of 2
        g[24:32] = np.array([i, 2], np.uint32).view(np.uint8)
    p = tmp_path / "two.nvdb"
    np.concatenate([a, b]).tofile(p)
    vf = I.import_nvdb(str(p))
    assert vf.data.nbytes == a.nbytes and vf.data[24:32].view(np.uint32).tolist() == [0, 1]
    assert vf.data[636:640].view(np.uint32)[0] == 1
    # errors: unknown magic, reversed endianness, truncated, BLOSC
    (tmp_path / "junk.nvdb").write_bytes(b"not a nanovdb file at all........")
    with pytest.raises(I.ImportError_) as e:
        I.import_nvdb(str(tmp_path / "junk.nvdb"))
    assert e.value.code == I.ERR_FORMAT and "unknown type" in str(e.value)
    (tmp_path / "swapped.nvdb").write_bytes(b"NanoVDB0"[::-1] + bytes(64))
    with pytest.raises(I.ImportError_) as e:
        I.import_nvdb(str(tmp_path / "swapped.nvdb"))
    assert "reversed endianness" in str(e.value)
    f = FIX["file/fog_float_zip.nvdb"].copy()
    f[14:16] = np.array([2], np.uint16).view(np.uint8)  # FileHeader::codec = BLOSC
    (tmp_path / "blosc.nvdb").write_bytes(f.tobytes())
    with pytest.raises(I.ImportError_) as e:
        I.import_nvdb(str(tmp_path / "blosc.nvdb"))
    assert e.value.code == I.ERR_UNSUPPORTED and "BLOSC" in str(e.value)
    (tmp_path / "short.nvdb").write_bytes(FIX["file/fog_float_none.nvdb"][:5000].tobytes())
    with pytest.raises(I.ImportError_):
        I.import_nvdb(str(tmp_path / "short.nvdb"))
    with pytest.raises(I.ImportError_) as e:
        I.import_nvdb(str(tmp_path / "missing.nvdb"))
    assert e.value.code == I.ERR_IO


# ------------------------------------------------------------------------------------------------- RAW
@pytest.mark.parametrize("fname,dtype,dvr", [
    ("bonsai_7x5x3_uint8.raw", np.uint8, capi.DVR_UFIXED8),
    ("scan_7x5x3_uint16.raw", np.uint16, capi.DVR_UFIXED16),
    ("scan_int16_7x5x3.raw", np.uint16, capi.DVR_UFIXED16),  # "int%i" selects the UNSIGNED type too
    ("field_7x5x3_float32.raw", np.float32, capi.DVR_FLOAT32),  # no int token: FLOAT32
    ("field_7x5x3.raw", np.float32, capi.DVR_FLOAT32),
    ("a_b_0x7x0x5x3_uint8.raw", np.uint8, capi.DVR_UFIXED8),  # %i reads hex: "0x7x0x5x3" -> 7 x 5 x 3
])
def test_import_raw_name_conventions(fname, dtype, dvr, tmp_path):
    rng = np.random.default_rng(1)
    vox = (rng.random((3, 5, 7)) * (200 if dtype != np.float32 else 1)).astype(dtype)
    p = tmp_path / fname
    vox.tofile(p)
    vf = I.import_raw(str(p))
    assert vf.kind == I.STRUCTURED and vf.data_type == dvr and vf.dims == (7, 5, 3)
    assert vf.origin == (0, 0, 0) and vf.spacing == (1, 1, 1) and vf.name == fname
    assert np.array_equal(vf.data, vox)
    scale = {np.uint8: 255.0, np.uint16: 65535.0, np.float32: 1.0}[dtype]
    assert vf.has_value_range
    np.testing.assert_allclose(vf.value_range, (vox.min() / scale, vox.max() / scale), rtol=1e-6)
    assert I.import_volume_file(str(p)).dims == (7, 5, 3)


def test_import_raw_errors(tmp_path):
    (tmp_path / "nodims_uint8.raw").write_bytes(bytes(10))
    with pytest.raises(I.ImportError_) as e:
        I.import_raw(str(tmp_path / "nodims_uint8.raw"))
    assert e.value.code == I.ERR_FORMAT and "unable to parse info" in str(e.value)
    (tmp_path / "short_4x4x4_uint8.raw").write_bytes(bytes(10))
    with pytest.raises(I.ImportError_) as e:
        I.import_raw(str(tmp_path / "short_4x4x4_uint8.raw"))
    assert e.value.code == I.ERR_IO
    (tmp_path / "wide_2x2x2_uint32.raw").write_bytes(bytes(32))
    with pytest.raises(I.ImportError_) as e:
        I.import_raw(str(tmp_path / "wide_2x2x2_uint32.raw"))
    assert e.value.code == I.ERR_UNSUPPORTED
    with pytest.raises(I.ImportError_):
        I.import_raw("relative_name_2x2x2.raw")  # fileOf() needs a directory separator, like the reference
    with pytest.raises(I.ImportError_) as e:
        I.import_volume_file(str(tmp_path / "thing.xyz"))
    assert "no loader for file type '.xyz'" in str(e.value)


# ------------------------------------------------------------------------------------------------- MHD
def test_import_mhd(tmp_path):
    rng = np.random.default_rng(3)
    vox = (rng.random((4, 6, 5)) * 60000).astype(np.uint16)
    vox.tofile(tmp_path / "ct.zraw")
    (tmp_path / "ct.mhd").write_text(
        "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\n"
        "DimSize = 5 6 4\nElementSpacing = 0.5 0.25 2.0\nElementType = MET_SHORT\nElementDataFile = ct.zraw\n")
    vf = I.import_mhd(str(tmp_path / "ct.mhd"))
    assert vf.dims == (5, 6, 4) and vf.data_type == capi.DVR_UFIXED16  # MET_SHORT -> UFIXED16, import_MHD.cpp:63
    assert np.array_equal(vf.data, vox)
    assert vf.spacing == (1, 1, 1) and vf.header_spacing == (0.5, 0.25, 2.0)  # parsed, not applied (reference)
    assert vf.name.endswith("/ct.zraw")
    np.testing.assert_allclose(vf.value_range, (vox.min() / 65535.0, vox.max() / 65535.0), rtol=1e-6)
    for met, dt, dvr in (("MET_UCHAR", np.uint8, capi.DVR_UFIXED8), ("MET_FLOAT", np.float32, capi.DVR_FLOAT32),
                         ("MET_DOUBLE", np.float64, capi.DVR_FLOAT64)):
        v = (rng.random((2, 3, 4)) * 100).astype(dt)
        v.tofile(tmp_path / "v.bin")
        (tmp_path / "v.mhd").write_text(f"DimSize = 4 3 2\nElementType = {met}\nElementDataFile = v.bin\n")
        got = I.import_volume_file(str(tmp_path / "v.mhd"))
        assert got.data_type == dvr and np.array_equal(got.data, v)
    (tmp_path / "bad.mhd").write_text("DimSize = 4 3 2\nElementType = MET_LONG\nElementDataFile = v.bin\n")
    with pytest.raises(I.ImportError_) as e:
        I.import_mhd(str(tmp_path / "bad.mhd"))
    assert e.value.code == I.ERR_UNSUPPORTED
    (tmp_path / "nodata.mhd").write_text("DimSize = 4 3 2\nElementType = MET_UCHAR\nElementDataFile = gone.bin\n")
    with pytest.raises(I.ImportError_) as e:
        I.import_mhd(str(tmp_path / "nodata.mhd"))
    assert e.value.code == I.ERR_IO


# ------------------------------------------------------------------------------------------------- VTI
def _vti(vox, fmt, compress=False, header="UInt32", appended_encoding="raw", extra_arrays=""):
    """A VTK XML ImageData file per the VTK file-formats document."""
    nz, ny, nx = vox.shape
    tname = {np.dtype(np.float32): "Float32", np.dtype(np.float64): "Float64", np.dtype(np.uint8): "UInt8",
             np.dtype(np.int8): "Int8", np.dtype(np.uint16): "UInt16", np.dtype(np.int16): "Int16"}[vox.dtype]
    hfmt = "<Q" if header == "UInt64" else "<I"
    raw = vox.tobytes()

    def block(b64):
        if not compress:
            data = struct.pack(hfmt, len(raw)) + raw
            return base64.b64encode(data) if b64 else data
        bs = 1000
        chunks = [raw[i:i + bs] for i in range(0, len(raw), bs)]
        comp = [zlib.compress(c) for c in chunks]
        last = len(chunks[-1]) if len(chunks[-1]) != bs else 0
        head = struct.pack(hfmt, len(chunks)) + struct.pack(hfmt, bs) + struct.pack(hfmt, last)
        head += b"".join(struct.pack(hfmt, len(c)) for c in comp)
        body = b"".join(comp)
        return (base64.b64encode(head) + base64.b64encode(body)) if b64 else head + body

    attrs = f'type="ImageData" version="1.0" byte_order="LittleEndian" header_type="{header}"'
    if compress:
        attrs += ' compressor="vtkZLibDataCompressor"'
    ext = f"0 {nx - 1} 0 {ny - 1} 0 {nz - 1}"
    out = [f'<?xml version="1.0"?>\n<VTKFile {attrs}>\n'
           f'  <ImageData WholeExtent="{ext}" Origin="-1.5 0.25 3" Spacing="0.5 0.125 2">\n'
           f'    <Piece Extent="{ext}">\n      <!-- point data -->\n      <PointData Scalars="density">\n{extra_arrays}']
    tail = b""
    if fmt == "ascii":
        vals = " ".join(repr(float(v)) if vox.dtype.kind == "f" else str(int(v)) for v in vox.ravel())
        out.append(f'        <DataArray type="{tname}" Name="density" format="ascii">\n{vals}\n        </DataArray>\n')
    elif fmt == "binary":
        out.append(f'        <DataArray type="{tname}" Name="density" format="binary">\n'
                   f'          {block(True).decode()}\n        </DataArray>\n')
    else:
        pad = b"\x07" * 12  # another array's bytes before ours: the offset attribute must be honoured
        first = struct.pack(hfmt, len(pad)) + pad
        if appended_encoding == "base64":
            first = base64.b64encode(first)
        out.append(f'        <DataArray type="{tname}" Name="density" format="appended" offset="{len(first)}"/>\n')
        tail = (f'  <AppendedData encoding="{appended_encoding}">\n   _').encode() + first + \
            block(appended_encoding == "base64") + b"\n  </AppendedData>\n"
    out.append('      </PointData>\n      <CellData/>\n    </Piece>\n  </ImageData>\n')
    return "".join(out).encode() + tail + b"</VTKFile>\n"


@pytest.mark.parametrize("fmt,compress,header,enc", [
    ("ascii", False, "UInt32", "raw"), ("binary", False, "UInt32", "raw"), ("binary", False, "UInt64", "raw"),
    ("binary", True, "UInt32", "raw"), ("binary", True, "UInt64", "raw"), ("appended", False, "UInt32", "raw"),
    ("appended", False, "UInt64", "base64"), ("appended", True, "UInt32", "raw"), ("appended", True, "UInt64", "base64")])
@pytest.mark.parametrize("dtype,dvr", [(np.float32, capi.DVR_FLOAT32), (np.uint16, capi.DVR_UFIXED16)])
def test_import_vti_every_encoding(fmt, compress, header, enc, dtype, dvr, tmp_path):
    rng = np.random.default_rng(4)
    vox = (rng.random((6, 9, 11)) * (1 if dtype == np.float32 else 65000)).astype(dtype)
    p = tmp_path / "img.vti"
    p.write_bytes(_vti(vox, fmt, compress, header, enc))
    vf = I.import_vti(str(p))
    assert vf.kind == I.STRUCTURED and vf.data_type == dvr and vf.dims == (11, 9, 6)
    assert vf.origin == (-1.5, 0.25, 3.0) and vf.spacing == (0.5, 0.125, 2.0)  # import_VTI.cpp:84-85
    assert np.array_equal(vf.data, vox)
    assert vf.name == "img.vti" and vf.has_value_range


def test_import_vti_skips_multi_component_arrays_and_reports_errors(tmp_path):
    vox = np.arange(24, dtype=np.uint8).reshape(2, 3, 4)
    vec = ('        <DataArray type="Float32" Name="velocity" NumberOfComponents="3" format="ascii">\n'
           + " ".join("0.5" for _ in range(72)) + '\n        </DataArray>\n')
    (tmp_path / "a.vti").write_bytes(_vti(vox, "binary", extra_arrays=vec))
    vf = I.import_volume_file(str(tmp_path / "a.vti"))
    assert vf.data_type == capi.DVR_UFIXED8 and np.array_equal(vf.data, vox)
    for dt in (np.int8, np.int16, np.float64):
        v = (np.arange(24) - 5).astype(dt).reshape(2, 3, 4)
        (tmp_path / "t.vti").write_bytes(_vti(v, "appended", True))
        assert np.array_equal(I.import_vti(str(tmp_path / "t.vti")).data, v)
    (tmp_path / "poly.vti").write_text('<VTKFile type="PolyData"><PolyData/></VTKFile>')
    with pytest.raises(I.ImportError_) as e:
        I.import_vti(str(tmp_path / "poly.vti"))
    assert e.value.code == I.ERR_FORMAT
    bad = _vti(vox.astype(np.float32), "binary").replace(b"Float32", b"Int32")
    (tmp_path / "i32.vti").write_bytes(bad)
    with pytest.raises(I.ImportError_) as e:
        I.import_vti(str(tmp_path / "i32.vti"))
    assert e.value.code == I.ERR_UNSUPPORTED
    trunc = _vti(vox.astype(np.float32), "appended")
    (tmp_path / "trunc.vti").write_bytes(trunc[:len(trunc) - 80])
    with pytest.raises(I.ImportError_):
        I.import_vti(str(tmp_path / "trunc.vti"))


def test_import_vti_rejects_block_sizes_that_overflow(tmp_path):
    """A crafted vtkZLibDataCompressor header whose (nblocks-1)*blockSize + lastSize wraps past 2^64 to a small value
    must be refused before any buffer is sized from it (it used to make uncompress() write past the heap block)."""
    vox = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    raw = vox.tobytes()
    comp = zlib.compress(raw)
    for nblocks, bs, last in ((3, 1 << 63, len(raw)), (2, (1 << 64) - 8, len(raw) + 8), (1, 0, 0), (2, 16, 32)):
        head = struct.pack("<QQQ", nblocks, bs, last) + b"".join(struct.pack("<Q", len(comp)) for _ in range(nblocks))
        body = comp * nblocks
        blob = (base64.b64encode(head) + base64.b64encode(body)).decode()
        xml = ('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" '
               'header_type="UInt64" compressor="vtkZLibDataCompressor">\n  <ImageData WholeExtent="0 3 0 2 0 1" '
               'Origin="0 0 0" Spacing="1 1 1">\n    <Piece Extent="0 3 0 2 0 1">\n      <PointData Scalars="d">\n'
               f'        <DataArray type="Float32" Name="d" format="binary">\n{blob}\n        </DataArray>\n'
               '      </PointData>\n    </Piece>\n  </ImageData>\n</VTKFile>\n')
        (tmp_path / "evil.vti").write_text(xml)
        with pytest.raises(I.ImportError_) as e:
            I.import_vti(str(tmp_path / "evil.vti"))
        assert e.value.code == I.ERR_FORMAT
    # compressed sizes larger than the payload that is present
    head = struct.pack("<QQQ", 1, len(raw), 0) + struct.pack("<Q", (1 << 64) - 1)
    blob = (base64.b64encode(head) + base64.b64encode(comp)).decode()
    (tmp_path / "evil2.vti").write_text(xml.replace(xml[xml.index('format="binary">') + 17:xml.index("\n        </DataArray>")], blob))
    with pytest.raises(I.ImportError_):
        I.import_vti(str(tmp_path / "evil2.vti"))


def test_import_nvdb_rejects_offsets_that_leave_the_grid(tmp_path):
    """Node offsets come from the file: a tree whose child offsets point outside the buffer is refused by the importer
    (before its host min/max walk) instead of being dereferenced."""
    g = nvdb_writer.fog_sphere(6.0)
    assert I.lib is not None
    root_off = 672 + int(g[672 + 24:672 + 32].view(np.int64)[0])
    ok = tmp_path / "ok.nvdb"
    g.tofile(ok)
    I.import_nvdb(str(ok))
    # (a) root tile count far beyond the buffer
    bad = g.copy()
    bad[root_off + 24:root_off + 28] = np.array([1 << 30], np.uint32).view(np.uint8)
    bad[20:24] = (bad[20:24].view(np.uint32) & ~np.uint32(4)).view(np.uint8)  # clear HasMinMax: force the walk
    (tmp_path / "tiles.nvdb").write_bytes(bad.tobytes())
    with pytest.raises(I.ImportError_) as e:
        I.import_nvdb(str(tmp_path / "tiles.nvdb"))
    assert e.value.code == I.ERR_FORMAT and "corrupt" in str(e.value)
    # (b) first tile's child offset past the end
    bad = g.copy()
    bad[root_off + 64 + 8:root_off + 64 + 16] = np.array([1 << 40], np.int64).view(np.uint8)
    (tmp_path / "child.nvdb").write_bytes(bad.tobytes())
    with pytest.raises(I.ImportError_) as e:
        I.import_nvdb(str(tmp_path / "child.nvdb"))
    assert e.value.code == I.ERR_FORMAT
    # (c) negative child offset
    bad = g.copy()
    bad[root_off + 64 + 8:root_off + 64 + 16] = np.array([-(1 << 33)], np.int64).view(np.uint8)
    (tmp_path / "neg.nvdb").write_bytes(bad.tobytes())
    with pytest.raises(I.ImportError_):
        I.import_nvdb(str(tmp_path / "neg.nvdb"))


def test_import_nvdb_holds_the_declared_grid_size_against_what_the_file_can_deliver(tmp_path):
    """The size of the grid allocation comes from the segment's FileMetaData.  A file that declares hundreds of
    gigabytes is refused before anything is allocated: Codec::NONE cannot hold more than the bytes that are left, a
    zlib stream cannot expand beyond 1032:1 (found by the mutation fuzzer, tools/fuzz_importers.cpp, under ASan)."""
    import resource
    fx = np.load(os.path.join(GOLD, "import_fixtures.npz"))
    for name, codec in (("fog_float_none", 0), ("fog_float_zip", 1)):
        blob = fx[f"file/{name}.nvdb"].copy()
        assert int(blob[14:16].view(np.uint16)[0]) == codec
        good = tmp_path / f"{name}.nvdb"
        blob.tofile(good)
        I.import_nvdb(str(good))
        bad = blob.copy()
        bad[16:24] = np.array([600 << 30], np.uint64).view(np.uint8)  # FileMetaData::gridSize of grid #0: 600 GiB
        (tmp_path / f"huge_{name}.nvdb").write_bytes(bad.tobytes())
        before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        with pytest.raises(I.ImportError_) as e:
            I.import_nvdb(str(tmp_path / f"huge_{name}.nvdb"))
        assert e.value.code in (I.ERR_FORMAT, I.ERR_IO)
        assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 256 * 1024  # KiB: nothing of that size was touched
        # a size only slightly too large for the stream is still caught by the decoder itself
        bad = blob.copy()
        bad[16:24] = (blob[16:24].view(np.uint64) + np.uint64(4096)).view(np.uint8)
        (tmp_path / f"off_{name}.nvdb").write_bytes(bad.tobytes())
        with pytest.raises(I.ImportError_):
            I.import_nvdb(str(tmp_path / f"off_{name}.nvdb"))


def test_path_helpers_keep_the_importers_conventions(tmp_path):
    """fileOf/extensionOf/splitString drive the RAW name parser: dims come from '_'-separated tokens of the base name."""
    v = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    d = tmp_path / "dir.with.dots"
    d.mkdir()
    p = d / "a__b_4x3x2_uint8.raw"  # consecutive delimiters give empty tokens, as std::getline does
    v.tofile(p)
    vf = I.import_raw(str(p))
    assert vf.dims == (4, 3, 2) and vf.name == "a__b_4x3x2_uint8.raw" and np.array_equal(vf.data, v)
    assert I.import_volume_file(str(p)).dims == (4, 3, 2)  # extension taken after the LAST dot of the path


# ------------------------------------------------------------------------------------------------- range
def test_compute_scalar_range_normalisation():
    assert I.compute_scalar_range(np.array([3, 200, 17], np.uint8), capi.DVR_UFIXED8) == (np.float32(3 / 255), np.float32(200 / 255))
    lo, hi = I.compute_scalar_range(np.array([-128, 5, 127], np.int8), capi.DVR_FIXED8)
    assert lo == -1.0 and hi == 1.0  # max(v/127, -1)
    lo, hi = I.compute_scalar_range(np.array([-32768, 32767], np.int16), capi.DVR_FIXED16)
    assert lo == -1.0 and hi == 1.0
    assert I.compute_scalar_range(np.array([65535, 0], np.uint16), capi.DVR_UFIXED16) == (0.0, 1.0)
    assert I.compute_scalar_range(np.array([2.5, -7.25], np.float32), capi.DVR_FLOAT32) == (-7.25, 2.5)
    assert I.compute_scalar_range(np.array([1e-3, 9.0], np.float64), capi.DVR_FLOAT64) == (np.float32(1e-3), 9.0)
    with pytest.raises(I.ImportError_) as e:
        I.compute_scalar_range(np.zeros(4, np.float16), capi.DVR_FLOAT16)
    assert e.value.code == I.ERR_UNSUPPORTED
