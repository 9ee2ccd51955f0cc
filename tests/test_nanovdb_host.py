"""CPU suite, part 5: the oracle's NanoVDB reader/sampler against sample values produced by the reference's
vendored NanoVDB 32.7.0 (grid->worldToIndexF + SampleFromVoxels<Accessor,1>, exactly the calls of
gpu/sampleSpatialField.h:80-109) — committed in tests/golden/nvdb_reference_samples.npz, regenerated live when
oracle/_ref/libref_host.so is present."""
import os

import numpy as np
import pytest

import oracle_binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("key", ["r20", "r12_vs025"])
def test_oracle_nanovdb_sampler_matches_reference_samples(key):
    fog = np.load(os.path.join(GOLD, "nvdb_fog_spheres.npz"))
    ref = np.load(os.path.join(GOLD, "nvdb_reference_samples.npz"))
    blob = np.ascontiguousarray(fog[key])
    got = ob.nvdb_sample_oracle(blob, ref[key + "_xyz"])
    want = ref[key + "_val"]
    assert (want != 0).mean() > 0.2  # the probe positions do hit the fog
    assert np.abs(got - want).max() <= 2.4e-7  # the reference host build does not contract a + w*(b-a)
    assert (got == want).mean() > 0.9


@pytest.mark.parametrize("key", ["fp4", "fp8", "fp16", "fpn"])
def test_oracle_quantised_grid_sampler_matches_reference_samples(key):
    """Fp4 / Fp8 / Fp16 / FpN grids quantised by NanoVDB itself (tests/golden/make_nvdb_fixtures.py); the reference
    values come from its own NanoGrid<FpX> accessors + SampleFromVoxels, dispatched like volumeIntegration.h:128-159."""
    blob = np.ascontiguousarray(np.load(os.path.join(GOLD, "nvdb_quant_spheres.npz"))[key])
    ref = np.load(os.path.join(GOLD, "nvdb_reference_samples.npz"))
    assert blob[636:640].view(np.uint32)[0] == ob.NVDB_GRID_TYPES[key]
    got = ob.nvdb_sample_oracle(blob, ref[key + "_xyz"])
    want = ref[key + "_val"]
    assert (want != 0).mean() > 0.2
    assert np.abs(got - want).max() <= 2.4e-7
    assert (got == want).mean() > 0.9


def test_quantised_fixtures_are_what_the_reference_generates():
    if not ob.have_ref_host():
        pytest.skip("oracle/_ref/libref_host.so not built")
    quant = np.load(os.path.join(GOLD, "nvdb_quant_spheres.npz"))
    ref = np.load(os.path.join(GOLD, "nvdb_reference_samples.npz"))
    for key in ("fp4", "fp8", "fp16", "fpn"):
        fresh = ob.nvdb_fog_sphere_typed(key, 14.0)
        if key != "fpn":  # NanoVDB leaves the slack of its worst-case FpN allocation uninitialised: bytes differ per run
            assert np.array_equal(fresh, quant[key]), key
        for blob in (fresh, np.ascontiguousarray(quant[key])):
            assert np.array_equal(ob.nvdb_sample_reference(blob, ref[key + "_xyz"]), ref[key + "_val"]), key


@pytest.mark.parametrize("codec,tol", [("fp4", 1.0 / 30 + 1e-6), ("fp8", 1.0 / 510 + 1e-6), ("fp16", 1e-5), ("fpn", 1e-3)])
def test_own_writer_quantised_codecs(codec, tol):
    """visrtx_b200.nvdb_writer with a quantised codec: the grid decodes (O-cpu; the real NanoVDB when present) to
    the float grid within half a quantum (fixed widths) / the requested tolerance (FpN)."""
    from visrtx_b200 import nvdb_writer as W
    rng = np.random.default_rng(17)
    xyz = (rng.random((3000, 3)) * 44 - 22).astype(np.float32)
    base = ob.nvdb_sample_oracle(W.fog_sphere(18.0), xyz)
    blob = W.fog_sphere(18.0, codec=codec, tolerance=1e-3)
    assert blob[636:640].view(np.uint32)[0] == W.GRID_TYPES[codec]
    got = ob.nvdb_sample_oracle(blob, xyz)
    assert np.abs(got - base).max() <= tol
    if codec != "fp16":
        assert np.abs(got - base).max() > 0  # it really is quantised
    if ob.have_ref_host():
        assert ob.nvdb_is_valid_reference(blob)
        assert np.abs(ob.nvdb_sample_reference(blob, xyz) - got).max() <= 2.4e-7
    if codec == "fpn":  # variable leaf sizes: homogeneous interior leaves take 1 bit
        assert blob.nbytes < W.fog_sphere(18.0, codec="fp8").nbytes * 1.3


def test_fixture_is_what_the_reference_generates():
    if not ob.have_ref_host():
        pytest.skip("oracle/_ref/libref_host.so not built")
    fog = np.load(os.path.join(GOLD, "nvdb_fog_spheres.npz"))
    assert np.array_equal(ob.nvdb_fog_sphere(20.0), fog["r20"])
    ref = np.load(os.path.join(GOLD, "nvdb_reference_samples.npz"))
    live = ob.nvdb_sample_reference(np.ascontiguousarray(fog["r20"]), ref["r20_xyz"])
    assert np.array_equal(live, ref["r20_val"])


def test_grid_header_fields():
    fog = np.load(os.path.join(GOLD, "nvdb_fog_spheres.npz"))
    blob = fog["r20"]
    assert blob[:8].tobytes() in (b"NanoVDB0", b"NanoVDB1")
    assert blob[636:640].view(np.uint32)[0] == 1  # GridType::Float
    wb = blob[560:608].view(np.float64)
    assert tuple(wb) == (-20.0, -20.0, -20.0, 21.0, 21.0, 21.0)


def test_own_nanovdb_writer_is_read_back_by_the_oracle_and_by_the_reference():
    """visrtx_b200.nvdb_writer output: sampled by O-cpu it equals dense trilinear interpolation of the source
    block; when the reference's NanoVDB is available (libref_host.so) the real library reads it back too."""
    import ctypes as C
    from visrtx_b200 import nvdb_writer as W
    rng = np.random.default_rng(11)
    dense = rng.random((21, 13, 30)).astype(np.float32)
    dense[dense < 0.4] = 0.0  # sparse: background voxels are not stored
    origin = (-9, 4, -17)
    blob = W.write_float_grid(dense, index_origin=origin, voxel_size=0.5, world_origin=(1.0, -2.0, 0.25))
    assert blob.nbytes % 32 == 0 and blob[:8].tobytes() == b"NanoVDB0"
    pts_idx = rng.random((4000, 3)) * (np.array(dense.shape) + 3) - 2 + np.array(origin)
    xyz = (pts_idx * 0.5 + np.array([1.0, -2.0, 0.25])).astype(np.float32)
    got = ob.nvdb_sample_oracle(blob, xyz)
    # dense trilinear ground truth in index space
    pi = (xyz.astype(np.float64) - np.array([1.0, -2.0, 0.25])) / 0.5
    i0 = np.floor(pi).astype(int)
    u = pi - i0
    want = np.zeros(len(xyz))
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (u[:, 0] if dx else 1 - u[:, 0]) * (u[:, 1] if dy else 1 - u[:, 1]) * (u[:, 2] if dz else 1 - u[:, 2])
                ii = i0 + np.array([dx, dy, dz]) - np.array(origin)
                ok = ((ii >= 0) & (ii < np.array(dense.shape))).all(axis=1)
                vals = np.where(ok, dense[np.clip(ii[:, 0], 0, 20), np.clip(ii[:, 1], 0, 12), np.clip(ii[:, 2], 0, 29)], 0.0)
                want += w * vals
    assert np.abs(got - want).max() < 2e-6
    if ob.have_ref_host():
        assert np.abs(ob.nvdb_sample_reference(blob, xyz) - got).max() <= 2.4e-7
        lib = ob.refhost()
        wb, vs, ib, gt, av = (C.c_double * 6)(), (C.c_double * 3)(), (C.c_int * 6)(), C.c_uint(), C.c_ulonglong()
        lib.refhost_nvdb_info(blob.ctypes.data_as(C.c_void_p), wb, vs, ib, C.byref(gt), C.byref(av))
        assert gt.value == 1 and av.value == int((dense != 0).sum()) and tuple(vs) == (0.5, 0.5, 0.5)
