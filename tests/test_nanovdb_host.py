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
