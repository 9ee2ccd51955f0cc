#!/bin/bash
# round 2, GPU call E (8 GPUs): multi-process parity at 2/4/8, C2 sort-last fused vs legacy, C4 at 2/4/8, C3 sort-first
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 420 python -m pytest tests/test_gpu_multigpu.py -q -m gpu > gpurun_out/r02e_pytest.log 2>&1
tail -5 gpurun_out/r02e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r02e_c2_n8_fused.json 2> gpurun_out/r02e_c2_n8_fused.err
tail -3 gpurun_out/r02e_c2_n8_fused.err
timeout 200 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 100 --warmup 10 --fused 0 --c4-scaling 0 > gpurun_out/r02e_c2_n8_legacy.json 2> gpurun_out/r02e_c2_n8_legacy.err
timeout 200 $TR --nproc-per-node 4 --master-port 29803 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r02e_c2_n4_fused.json 2> gpurun_out/r02e_c2_n4_fused.err
for n in 2 4 8; do
  timeout 240 $TR --nproc-per-node $n --master-port $((29810+n)) bench.py --gpus $n --steps 50 --warmup 5 --config c3 --mode sort-first --c4-scaling 0 > gpurun_out/r02e_c3_sf_n$n.json 2> gpurun_out/r02e_c3_sf_n$n.err
done
python - <<'PY'
import json
for f in ["r02e_c2_n8_fused", "r02e_c2_n8_legacy", "r02e_c2_n4_fused", "r02e_c3_sf_n2", "r02e_c3_sf_n4", "r02e_c3_sf_n8"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d.get("e2e", {}).get("value"), 1), "march_us", d["extra"].get("march_us"),
              "exchange_us", d["extra"].get("exchange_us"))
        print("   phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
        if "c4_scaling" in d["extra"]:
            print("   c4", json.dumps(d["extra"]["c4_scaling"])[:1500])
    except Exception as e:
        print(f, "ERR", e)
PY
