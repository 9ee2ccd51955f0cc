#!/bin/bash
# round 2, GPU call AB (4 GPUs): slab cuts balanced on the march PHASE inside the fused frame (SortLast.calibrate,
# fused rounds) — multi-process + ANARI parity at 2 and 4, then the N = 4 bench line as the driver runs it
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_anari_multigpu.py -q -x -k "not 8" ) > gpurun_out/r02ab_pytest.log 2>&1
tail -3 gpurun_out/r02ab_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 300 $TR --master-port 29871 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02ab_n4.json 2> gpurun_out/r02ab_n4.err
tail -4 gpurun_out/r02ab_n4.err
python - <<'PY'
import json
f = "gpurun_out/r02ab_n4.json"
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    x = d["extra"]
    ph = x.get("fused_phases_us_per_rank", {}).get("ranks")
    print("fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "us/frame", round(1e3 * d["ms_per_step"], 1), "alone", x.get("march_alone_us_per_rank"))
    print("march inside", [p[0] for p in ph], "total", [p[4] for p in ph], "parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k not in ("what", "tolerance")})
    b = x.get("slab_balance") or {}
    print("balance", b.get("initial"), "->", b.get("final"))
    for r in b.get("rounds", []):
        print("   ", r)
    c = x.get("c3_sort_first")
    print("c3_sort_first", c if isinstance(c, str) else {k: c[k] for k in ("value", "single_gpu_value", "speedup_over_one_gpu", "tile_band", "ms_per_step_by_tile_band")}, None if isinstance(c, str) else c["parity_vs_single"]["bit_identical"])
except Exception as e:
    print(f, "ERR", e)
PY
