#!/bin/bash
# round 2, GPU call Y (2 GPUs): which part of the fused launch stretches the march phase (505 us inside, 470 us alone at
# N = 2)?  DVR_B200_SLAB_DEBUG: 1 no region flags, 2 no background strip, 4 no compositing (frames incomplete: timing only)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29840
for dbg in 7 5 6 4 0; do
  port=$((port+1))
  DVR_B200_SLAB_DEBUG=$dbg timeout 120 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --c3-sort-first 0 --no-cpu-baseline > gpurun_out/r02y_n2_dbg$dbg.json 2> gpurun_out/r02y_n2_dbg$dbg.err
done
python - <<'PY'
import json
for dbg in [7, 5, 6, 4, 0]:
    f = f"gpurun_out/r02y_n2_dbg{dbg}.json"
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        ph = x.get("fused_phases_us_per_rank", {}).get("ranks")
        print("dbg", dbg, "fps", round(d["value"], 1), "us/frame", round(1e3 * d["ms_per_step"], 1), "alone", x.get("march_alone_us_per_rank"),
              "march inside", [p[0] for p in ph], "total", [p[4] for p in ph])
    except Exception as e:
        print(f, "ERR", e)
PY
