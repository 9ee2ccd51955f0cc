#!/bin/bash
# round 2, GPU call X (2 GPUs): the N = 2 bench line as the driver runs it (now with C3 sort-first inside), multi-process
# parity at 2, and the poll interval of the warps that wait for region flags (first interval / doubling cap)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 300 $TR --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02x_n2.json 2> gpurun_out/r02x_n2.err
tail -4 gpurun_out/r02x_n2.err
( timeout 300 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_anari_multigpu.py -q -x -k "2 or anari" ) > gpurun_out/r02x_pytest.log 2>&1
tail -3 gpurun_out/r02x_pytest.log
port=29820
for cfg in "100 0" "100 1600" "400 3200" "1000 0" "2000 0"; do
  set -- $cfg; port=$((port+1))
  DVR_B200_SPIN_NS=$1 DVR_B200_SPIN_CAP_NS=$2 timeout 120 $TR --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --c3-sort-first 0 --no-cpu-baseline > gpurun_out/r02x_n2_spin$1_$2.json 2> gpurun_out/r02x_n2_spin$1_$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02x_n2*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        x = d["extra"]
        ph = x.get("fused_phases_us_per_rank", {}).get("ranks")
        print(f, "fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "march_alone", x.get("march_alone_us_per_rank"),
              "exchange_us", round(x.get("exchange_us", 0), 1), "parity", (d.get("parity_vs_single") or {}).get("pass"))
        print("    phases", ph)
        if "c3_sort_first" in x:
            c = x["c3_sort_first"]
            print("    c3_sort_first", c if isinstance(c, str) else {k: c[k] for k in ("value", "single_gpu_value", "speedup_over_one_gpu", "tile_band", "ms_per_step_by_tile_band", "setup_s")}, c if isinstance(c, str) else c["parity_vs_single"]["bit_identical"])
    except Exception as e:
        print(f, "ERR", e)
PY
