"""Generates the NanoVDB fixtures of tests/golden/ IN THE BUILD CONTAINER (needs oracle/_ref/libref_host.so, i.e.
the reference tree with its vendored NanoVDB 32.7.0):

    python tests/golden/make_nvdb_fixtures.py

  nvdb_fog_spheres.npz        float fog spheres from nanovdb::tools::createFogVolumeSphere<float>
  nvdb_quant_spheres.npz      the same sphere quantised by NanoVDB itself: Fp4 / Fp8 / Fp16 / FpN grids
  nvdb_reference_samples.npz  values of the reference's sampler (grid->worldToIndexF + SampleFromVoxels<Acc,1>,
                              gpu/sampleSpatialField.h:80-109) at seeded random world positions, per grid
The generation is deterministic; the CPU suite checks the committed files against a live regeneration when the
reference library is present (tests/test_nanovdb_host.py).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import oracle_binding as ob  # noqa: E402

QUANT_RADIUS = 14.0


def probe_positions(seed, n, extent):
    rng = np.random.default_rng(seed)
    return ((rng.random((n, 3)) * 2.0 - 1.0) * extent).astype(np.float32)


def main():
    fog = {"r20": ob.nvdb_fog_sphere(20.0), "r12_vs025": ob.nvdb_fog_sphere(12.0, voxel_size=0.25, center=(1.0, -0.5, 2.0))}
    old = np.load(os.path.join(HERE, "nvdb_reference_samples.npz"))
    samples = {}
    for key, blob in fog.items():
        xyz = old[key + "_xyz"]  # keep the committed probe positions
        samples[key + "_xyz"] = xyz
        samples[key + "_val"] = ob.nvdb_sample_reference(blob, xyz)
    quant = {}
    for i, t in enumerate(("fp4", "fp8", "fp16", "fpn")):
        blob = ob.nvdb_fog_sphere_typed(t, QUANT_RADIUS)
        quant[t] = blob
        xyz = probe_positions(100 + i, 5000, QUANT_RADIUS + 4.0)
        samples[t + "_xyz"] = xyz
        samples[t + "_val"] = ob.nvdb_sample_reference(blob, xyz)
    np.savez_compressed(os.path.join(HERE, "nvdb_fog_spheres.npz"), **fog)
    np.savez_compressed(os.path.join(HERE, "nvdb_quant_spheres.npz"), **quant)
    np.savez_compressed(os.path.join(HERE, "nvdb_reference_samples.npz"), **samples)
    for k, v in {**fog, **quant}.items():
        print(k, v.nbytes, "bytes, grid type", int(v[636:640].view(np.uint32)[0]))


if __name__ == "__main__":
    main()
