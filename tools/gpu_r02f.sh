#!/bin/bash
# round 2, GPU call F (2 GPUs): where does the march phase of the fused launch lose time?
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29900
for dbg in 0 1 2 4 6 7; do
  port=$((port+1))
  DVR_B200_SLAB_DEBUG=$dbg timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --c4-scaling 0 > gpurun_out/r02f_n2_dbg$dbg.json 2> gpurun_out/r02f_n2_dbg$dbg.err
done
python - <<'PY'
import json
for dbg in [0, 1, 2, 4, 6, 7]:
    f = f"r02f_n2_dbg{dbg}"
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "fps", round(d.get("value"), 1), "alone", d["extra"].get("march_alone_us_per_rank"), "phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"))
    except Exception as e:
        print(f, "ERR", e)
PY
