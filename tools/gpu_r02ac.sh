#!/bin/bash
# round 2, GPU call AC (2 GPUs): the N = 2 bench line as the driver runs it, now with extra.anari_multi_gpu (rank 0 drives
# both GPUs through the ANARI C API in one process) after extra.c3_sort_first
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 200 $TR --master-port 29881 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02ac_n2.json 2> gpurun_out/r02ac_n2.err
tail -5 gpurun_out/r02ac_n2.err
python - <<'PY'
import json
f = "gpurun_out/r02ac_n2.json"
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    x = d["extra"]
    print("fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "parity", (d.get("parity_vs_single") or {}).get("pass"))
    print("anari_multi_gpu", x.get("anari_multi_gpu"))
    c = x.get("c3_sort_first")
    print("c3_sort_first", c if isinstance(c, str) else (c["value"], c["speedup_over_one_gpu"], c["parity_vs_single"]["bit_identical"]))
    print("balance rounds", [(r.get("march_ms") or r.get("fused_march_ms")) for r in (x.get("slab_balance") or {}).get("rounds", [])])
except Exception as e:
    print(f, "ERR", e)
PY
