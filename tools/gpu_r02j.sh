#!/bin/bash
# round 2, GPU call J (1 GPU): mixed scenes incl. the ANARI objects, then the ANARI suite (regression)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_anari.py -m gpu -q -x -s ) > gpurun_out/r02j_pytest.log 2>&1
tail -40 gpurun_out/r02j_pytest.log
