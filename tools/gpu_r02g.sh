#!/bin/bash
# round 2, GPU call G (2 GPUs): interleaved fused launch — parity, then timing (full / no composite between tiles)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multigpu.py -q -m gpu -x > gpurun_out/r02g_pytest.log 2>&1
tail -4 gpurun_out/r02g_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29950
for dbg in 0 8; do
  port=$((port+1))
  DVR_B200_SLAB_DEBUG=$dbg timeout 120 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --c4-scaling 0 > gpurun_out/r02g_n2_dbg$dbg.json 2> gpurun_out/r02g_n2_dbg$dbg.err
done
python - <<'PY'
import json
for dbg in [0, 8]:
    f = f"r02g_n2_dbg{dbg}"
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "fps", round(d.get("value"), 1), "alone", d["extra"].get("march_alone_us_per_rank"), "phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k not in ("what", "tolerance")})
    except Exception as e:
        print(f, "ERR", e)
PY
