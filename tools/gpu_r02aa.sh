#!/bin/bash
# round 2, GPU call AA (1 GPU): end-of-round state check — whole GPU suite, smoke, both bench arms as the driver runs
# them, launch list of the bench and one ncu --set full capture of the C2 frame kernel (round-2 build)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02aa_pytest.log 2>&1
tail -4 gpurun_out/r02aa_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r02aa_bench_ref.json 2> gpurun_out/r02aa_bench_ref.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err
tail -4 gpurun_out/r02aa_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02aa_launches.csv python bench.py --steps 2 --warmup 3 --extra 0 --no-cpu-baseline --extra-configs "" > gpurun_out/r02aa_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dvrFrameKernel -s 6 -c 1 -f -o gpurun_out/r02aa_c2 python tools/profile_scene.py --what c2 --frames 10 > gpurun_out/r02aa_c2_ncu.log 2>&1
ncu -i gpurun_out/r02aa_c2.ncu-rep --page raw --csv > gpurun_out/r02aa_c2_ncu.csv 2>/dev/null
python - <<'PY'
import json
for f in ["r02aa_bench_ref", "r02aa_bench"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), (d.get("parity") or {}).get("pass"))
        x = d.get("extra", {})
        for k, v in (x.get("configs") or {}).items():
            print("  ", k, v.get("value") if isinstance(v, dict) else v, (v.get("parity") or {}).get("pass") if isinstance(v, dict) else "")
        if "dpt" in x:
            print("   dpt", x["dpt"])
        if "time_varying" in x:
            print("   tv", x["time_varying"].get("fps") if isinstance(x["time_varying"], dict) else x["time_varying"])
    except Exception as e:
        print(f, "ERR", e)
PY
