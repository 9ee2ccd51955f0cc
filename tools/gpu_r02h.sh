#!/bin/bash
# round 2, GPU call H (1 GPU): state check after the container was re-created — whole GPU suite, smoke, both bench
# arms exactly as the driver runs them, kernel launch list of the bench
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02h_pytest.log 2>&1
tail -8 gpurun_out/r02h_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
tail -6 gpurun_out/r02h_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --extra 0 --no-cpu-baseline --extra-configs "" > gpurun_out/r02h_ncu_bench.log 2>&1
python - <<'PY'
import json
for f in ["r02h_bench_ref", "r02h_bench"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("parity"))
        for k, v in (d.get("extra", {}).get("configs") or {}).items():
            print("  ", k, v.get("value"), v.get("e2e", {}).get("value") if isinstance(v.get("e2e"), dict) else v.get("e2e"), (v.get("roofline") or {}).get("frac"), v.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
