#!/bin/bash
# round 2, GPU call A: goldens for the new scenes, the whole GPU suite, both bench arms, L2 fetch granularity A/B
set -x
mkdir -p gpurun_out
python tests/golden/make_golden.py --add bgimage_float3_default_spp2_f2,bgimage_srgb4_raycast_far,bgimage_u16x2_checkerboard_p5 > gpurun_out/r02a_golden.log 2>&1
cp gpurun_out/golden/refgpu_scenes.npz tests/golden/refgpu_scenes.npz
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multigpu.py > gpurun_out/r02a_pytest.log 2>&1
tail -30 gpurun_out/r02a_pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_bench.err
for g in 32 64 128; do
  DVR_B200_L2_FETCH=$g timeout 300 python bench.py --steps 30 --warmup 5 --extra 0 --no-cpu-baseline > gpurun_out/r02a_l2fetch_$g.json 2> gpurun_out/r02a_l2fetch_$g.err
done
python - <<'PY'
import json
for f in ["r02a_bench_ref", "r02a_bench", "r02a_l2fetch_32", "r02a_l2fetch_64", "r02a_l2fetch_128"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
