#!/bin/bash
# round 2, GPU call N (2 GPUs): held-ticket composite queue — parity at 2, bench N = 2 (interleaved / composite after the tiles)
mkdir -p gpurun_out
( time timeout 400 python -m pytest "tests/test_gpu_multigpu.py::test_multi_process_sort_first_and_sort_last[2]" \
   "tests/test_gpu_anari_multigpu.py::test_sort_last_through_anari_matches_the_single_gpu_frame[2]" -q -m gpu ) > gpurun_out/r02n_pytest.log 2>&1
tail -6 gpurun_out/r02n_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29900
for dbg in 0 8 16; do
  port=$((port+1))
  DVR_B200_SLAB_DEBUG=$dbg timeout 120 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --c4-scaling 0 --no-cpu-baseline > gpurun_out/r02n_n2_dbg$dbg.json 2> gpurun_out/r02n_n2_dbg$dbg.err
done
python - <<'PY'
import json
for dbg in [0, 8, 16]:
    f = f"r02n_n2_dbg{dbg}"
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "fps", round(d.get("value"), 1), "e2e", round(d["e2e"]["value"], 1), "alone", d["extra"].get("march_alone_us_per_rank"), "phases", d["extra"].get("fused_phases_us_per_rank", {}).get("ranks"))
        print("   parity", {k: v for k, v in (d.get("parity_vs_single") or {}).items() if k != "what" and k != "tolerance"})
    except Exception as e:
        print(f, "ERR", e)
PY
