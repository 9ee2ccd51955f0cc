#!/bin/bash
# round 2, GPU call O (8 GPUs): held-ticket composite queue at N = 8 / 4 — ANARI sort-last parity at 8, bench as the
# driver runs it (C4 scaling inside the N = 8 line), interleaved vs composite-after-tiles A/B
mkdir -p gpurun_out
( time timeout 300 python -m pytest "tests/test_gpu_multigpu.py::test_multi_process_sort_first_and_sort_last[8]" \
  "tests/test_gpu_anari_multigpu.py::test_sort_last_through_anari_matches_the_single_gpu_frame[8]" -q -m gpu ) > gpurun_out/r02o_pytest.log 2>&1
tail -4 gpurun_out/r02o_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02o_c2_n8.json 2> gpurun_out/r02o_c2_n8.err
tail -3 gpurun_out/r02o_c2_n8.err
DVR_B200_SLAB_DEBUG=8 timeout 200 $TR --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 --steps 50 --warmup 5 --c4-scaling 0 --no-cpu-baseline > gpurun_out/r02o_c2_n8_dbg8.json 2> gpurun_out/r02o_c2_n8_dbg8.err
( time timeout 300 $TR --nproc-per-node 4 --master-port 29802 bench.py --gpus 4 --steps 20 --warmup 5 ) > gpurun_out/r02o_c2_n4.json 2> gpurun_out/r02o_c2_n4.err
python - <<'PY'
import json
for f in ["r02o_c2_n8", "r02o_c2_n8_dbg8", "r02o_c2_n4"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d.get("e2e", {}).get("value"), 1), "march_us", x.get("march_us"),
              "exchange_us", x.get("exchange_us"), "alone", x.get("march_alone_us_per_rank"))
        print("   phases", x.get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
        if "c4_scaling" in x:
            c = x["c4_scaling"]
            print("   c4", {k: (round(v["value"], 1), {kk: vv for kk, vv in v.items() if kk.startswith("parity")}) for k, v in c["runs"].items()}, c.get("speedup_8_over_2"))
    except Exception as e:
        print(f, "ERR", e)
PY
