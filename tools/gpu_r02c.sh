#!/bin/bash
# round 2, GPU call C (2 GPUs): multi-process parity of the fused sort-last frame + A/B against the round-1 sequence
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q -m gpu -x > gpurun_out/r02c_pytest.log 2>&1
tail -25 gpurun_out/r02c_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02c_n2_fused.json 2> gpurun_out/r02c_n2_fused.err
timeout 400 $TR --master-port 29612 bench.py --gpus 2 --steps 50 --warmup 5 --fused 0 > gpurun_out/r02c_n2_legacy.json 2> gpurun_out/r02c_n2_legacy.err
tail -5 gpurun_out/r02c_n2_fused.err
python - <<'PY'
import json
for f in ["r02c_n2_fused", "r02c_n2_legacy"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "fps", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"), "march_us", d["extra"].get("march_us"),
              "exchange_us", d["extra"].get("exchange_us"), "parity", d.get("parity_vs_single"))
    except Exception as e:
        print(f, "ERR", e)
PY
