#!/bin/bash
# round 2, GPU call L (2 GPUs): per-region ticket claims in the fused sort-last kernel — parity tests at 2, bench N = 2
mkdir -p gpurun_out
( time timeout 400 python -m pytest "tests/test_gpu_multigpu.py::test_multi_process_sort_first_and_sort_last[2]" \
   tests/test_gpu_anari_multigpu.py -q -m gpu ) > gpurun_out/r02l_pytest.log 2>&1
tail -6 gpurun_out/r02l_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 $TR --nproc-per-node 2 --master-port 29802 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02l_c2_n2.json 2> gpurun_out/r02l_c2_n2.err
tail -4 gpurun_out/r02l_c2_n2.err
python - <<'PY'
import json
for f in ["r02l_c2_n2"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d.get("e2e", {}).get("value"), 1), "march_us", x.get("march_us"),
              "exchange_us", x.get("exchange_us"), "alone", x.get("march_alone_us_per_rank"))
        print("   phases", x.get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
    except Exception as e:
        print(f, "ERR", e)
PY
