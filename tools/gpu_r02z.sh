#!/bin/bash
# round 2, GPU call Z (2 GPUs): background strip as its own launch ahead of the fused frame (A/B against the in-kernel
# placement), multi-process + ANARI multi-GPU parity at 2 with it
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_anari_multigpu.py -q -x -k "2 or anari" ) > gpurun_out/r02z_pytest.log 2>&1
tail -3 gpurun_out/r02z_pytest.log
timeout 120 $TR --master-port 29861 bench.py --gpus 2 --steps 50 --warmup 5 --c3-sort-first 0 --no-cpu-baseline > gpurun_out/r02z_n2_bgkernel.json 2> gpurun_out/r02z_n2_bgkernel.err
DVR_B200_SLAB_BG_INLINE=1 timeout 120 $TR --master-port 29862 bench.py --gpus 2 --steps 50 --warmup 5 --c3-sort-first 0 --no-cpu-baseline > gpurun_out/r02z_n2_bginline.json 2> gpurun_out/r02z_n2_bginline.err
python - <<'PY'
import json
for v in ["bgkernel", "bginline"]:
    f = f"gpurun_out/r02z_n2_{v}.json"
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        ph = x.get("fused_phases_us_per_rank", {}).get("ranks")
        print(v, "fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "us/frame", round(1e3 * d["ms_per_step"], 1), "alone", x.get("march_alone_us_per_rank"),
              "march inside", [p[0] for p in ph], "total", [p[4] for p in ph], "launches", d.get("gpu_launches"), "parity", (d.get("parity_vs_single") or {}).get("pass"))
    except Exception as e:
        print(f, "ERR", e)
PY
