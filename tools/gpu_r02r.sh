#!/bin/bash
# round 2, GPU call R (2 GPUs): ownership-shift balancing (worker test) + bench N = 2 with calibration
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py > gpurun_out/r02r_worker.log 2>&1
grep -E "^\[sort-last x2 (balance|fused, first)|MGPU|Error|error" gpurun_out/r02r_worker.log | head -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29901 bench.py --gpus 2 --steps 30 --warmup 5 --c4-scaling 0 --no-cpu-baseline > gpurun_out/r02r_n2.json 2> gpurun_out/r02r_n2.err
tail -3 gpurun_out/r02r_n2.err
python - <<'PY'
import json
for f in ["r02r_n2"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d["e2e"]["value"], 1), "alone", x.get("march_alone_us_per_rank"), "phases", x.get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   balance", json.dumps(x.get("slab_balance"))[:900])
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
    except Exception as e:
        print(f, "ERR", e)
PY
