#!/bin/bash
# round 2, GPU call Q (8 GPUs): why is the march inside the fused launch slower than the march alone (rank 0: +36 us)?
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29900
for dbg in 8 10 12 14 9; do
  port=$((port+1))
  DVR_B200_SLAB_DEBUG=$dbg timeout 120 $TR --master-port $port bench.py --gpus 8 --steps 50 --warmup 5 --c4-scaling 0 --no-cpu-baseline > gpurun_out/r02q_n8_dbg$dbg.json 2> gpurun_out/r02q_n8_dbg$dbg.err
done
python - <<'PY'
import json
for dbg in [8, 10, 12, 14, 9]:
    f = f"r02q_n8_dbg{dbg}"
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        ph = d["extra"].get("fused_phases_us_per_rank", {}).get("ranks")
        print(f, "fps", round(d.get("value"), 1), "alone", d["extra"].get("march_alone_us_per_rank"))
        print("    march", [p[0] for p in ph], "total", [p[4] for p in ph])
    except Exception as e:
        print(f, "ERR", e)
PY
