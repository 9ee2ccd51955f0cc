#!/bin/bash
# round 2, GPU call M (2 GPUs): where does the interleaved fused launch lose time? DVR_B200_SLAB_DEBUG A/B
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29900
for dbg in 0 8 16 24 2 4 6; do
  port=$((port+1))
  DVR_B200_SLAB_DEBUG=$dbg timeout 120 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --c4-scaling 0 --no-cpu-baseline > gpurun_out/r02m_n2_dbg$dbg.json 2> gpurun_out/r02m_n2_dbg$dbg.err
done
python - <<'PY'
import json
for dbg in [0, 8, 16, 24, 2, 4, 6]:
    f = f"r02m_n2_dbg{dbg}"
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "fps", round(d.get("value"), 1), "alone", d["extra"].get("march_alone_us_per_rank"), "phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"))
    except Exception as e:
        print(f, "ERR", e)
PY
