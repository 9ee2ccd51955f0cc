"""CPU suite: the host-side pieces of bench.py's contract — presets, byte model, the committed DRAM-traffic figure,
and the loud failure without a GPU (the measured path has no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(*argv):
    old = sys.argv
    sys.argv = ["bench.py", *argv]
    try:
        return bench.parse_args()
    finally:
        sys.argv = old


def test_default_is_baseline_config_c2_on_one_gpu():
    a = _args()
    assert (a.gpus, a.size, a.width, a.height, a.rate, a.field, a.config) == (1, 1024, 1920, 1080, 0.5, "ml", "c2")
    assert a.warmup >= 3 and a.steps >= 50 and a.skip == 0
    name = bench.workload_name(a)
    assert name.startswith("C2: 1024^3 f32") and "1920x1080" in name and "default renderer" in name


@pytest.mark.parametrize("cfg,size,res,field,skip", [("c3", 2048, (3840, 2160), "shells", 1), ("c4", 4096, (1920, 1080), "ml", 0),
                                                      ("c5", 401, (1920, 1080), "fog", 1)])
def test_presets(cfg, size, res, field, skip):
    a = _args("--config", cfg)
    assert (a.size, (a.width, a.height), a.field, a.skip) == (size, res, field, skip)
    if cfg == "c4":
        assert a.mode == "sort-last"
    assert _args("--config", cfg, "--skip", "0").skip == 0  # an explicit --skip wins over the preset


def test_byte_model_and_committed_traffic():
    a = _args()
    cells = 261232
    b, model = bench.bytes_per_frame(a, cells, samples=76_000_000)
    assert b == cells * 4096 * 4 + 1920 * 1080 * 44 + 4096 and "macrocells" in model
    b2, model2 = bench.bytes_per_frame(a, cells, samples=1_000_000)  # rays sparser than macrocells: sector cap
    assert b2 == 1_000_000 * 128 + 1920 * 1080 * 44 + 4096 and "sector cap" in model2
    t = bench.measured_traffic(a, "single")
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        entry = json.load(f)[bench.traffic_key(a, "single")]
    assert t == entry["dram_bytes_per_launch"] and 4.0e9 < t < 5.5e9
    assert os.path.exists(os.path.join(ROOT, entry["source"].split(" ")[0]))
    assert bench.measured_traffic(_args("--config", "c3"), "single") is None  # no capture committed for it


def test_bench_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    for extra in ([], ["--impl", "reference"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", *extra],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode != 0
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert "no CUDA device" in line["error"]


def test_secondary_measurements_of_the_multi_gpu_lines_default_to_auto_and_can_be_switched_off():
    """N > 1 lines carry C3 sort-first, the ANARI one-process multi-GPU e2e and (at 8) C4 scaling as `extra.*`; each has
    an auto / on / off switch, and each is wrapped so that it can never take the headline line down."""
    a = _args()
    assert (a.c3_sort_first, a.anari_multi_gpu, a.c4_scaling) == (-1, -1, -1)
    b = _args("--c3-sort-first", "0", "--anari-multi-gpu", "0", "--c4-scaling", "0")
    assert (b.c3_sort_first, b.anari_multi_gpu, b.c4_scaling) == (0, 0, 0)
    import inspect
    src = inspect.getsource(bench.run_ours)
    assert 'out["extra"]["c3_sort_first"]' in src and 'out["extra"]["anari_multi_gpu"]' in src
    assert "except Exception" in inspect.getsource(bench.measure_anari_multi_gpu)
    assert 'backend="gloo"' in inspect.getsource(bench.measure_anari_multi_gpu)  # waiting ranks must not spin on their GPU
