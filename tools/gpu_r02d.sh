#!/bin/bash
# round 2, GPU call D (2 GPUs): phases of the fused sort-last launch, spin back-off sweep
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29700
for ns in 100; do
  port=$((port+1))
  DVR_B200_SPIN_NS=$ns timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02d_n2_spin$ns.json 2> gpurun_out/r02d_n2_spin$ns.err
done
python - <<'PY'
import json
for f in ["r02d_n2_spin100", "r02d_n2_spin500", "r02d_n2_spin2000"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "fps", round(d.get("value"), 1), "phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"), "march_us", d["extra"].get("march_us"))
    except Exception as e:
        print(f, "ERR", e)
PY
