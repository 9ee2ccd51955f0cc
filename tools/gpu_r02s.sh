#!/bin/bash
# round 2, GPU call S (8 GPUs): the bench exactly as the driver runs it at N = 8 (held tickets, calibrated cuts, C4
# scaling on 2/4/8 of the ranks) and at N = 4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 400 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02s_c2_n8.json 2> gpurun_out/r02s_c2_n8.err
tail -3 gpurun_out/r02s_c2_n8.err
( time timeout 200 $TR --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02s_c2_n4.json 2> gpurun_out/r02s_c2_n4.err
python - <<'PY'
import json
for f in ["r02s_c2_n8", "r02s_c2_n4"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d.get("e2e", {}).get("value"), 1), "march_us", x.get("march_us"),
              "exchange_us", x.get("exchange_us"), "alone", x.get("march_alone_us_per_rank"))
        print("   phases", x.get("fused_phases_us_per_rank", {}).get("ranks"))
        b = x.get("slab_balance") or {}
        print("   balance", b.get("initial"), "->", b.get("final"), [r["march_ms"] for r in b.get("rounds", [])])
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
        if "c4_scaling" in x:
            c = x["c4_scaling"]
            print("   c4", {k: (round(v["value"], 1), {kk: {a: b for a, b in vv.items() if a not in ("what", "tolerance")} for kk, vv in v.items() if kk.startswith("parity")}) for k, v in c["runs"].items()}, c.get("speedup_8_over_2"))
    except Exception as e:
        print(f, "ERR", e)
PY
