#!/bin/bash
# round 2, GPU call K (8 GPUs): multi-process + ANARI multi-GPU parity at 8, then the bench exactly as the driver runs
# it at N = 8 (C2 sort-last fused, C4 scaling on 2/4/8 of the ranks inside) and at N = 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 400 python -m pytest "tests/test_gpu_multigpu.py::test_multi_process_sort_first_and_sort_last[8]" \
   "tests/test_gpu_anari_multigpu.py::test_sort_last_through_anari_matches_the_single_gpu_frame[8]" \
   tests/test_gpu_anari_multigpu.py::test_sort_first_through_anari_is_bit_identical \
   tests/test_gpu_anari_multigpu.py::test_scenes_outside_the_distributed_paths_fall_back_to_the_display_gpu -q -m gpu ) > gpurun_out/r02k_pytest.log 2>&1
tail -6 gpurun_out/r02k_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 900 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02k_c2_n8.json 2> gpurun_out/r02k_c2_n8.err
tail -4 gpurun_out/r02k_c2_n8.err
( time timeout 300 $TR --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02k_c2_n4.json 2> gpurun_out/r02k_c2_n4.err
tail -4 gpurun_out/r02k_c2_n4.err
python - <<'PY'
import json
for f in ["r02k_c2_n8", "r02k_c2_n4"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d.get("e2e", {}).get("value"), 1), "march_us", x.get("march_us"),
              "exchange_us", x.get("exchange_us"), "alone", x.get("march_alone_us_per_rank"))
        print("   phases", x.get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
        if "c4_scaling" in x:
            print("   c4", json.dumps(x["c4_scaling"])[:2500])
    except Exception as e:
        print(f, "ERR", e)
PY
