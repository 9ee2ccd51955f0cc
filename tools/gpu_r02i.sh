#!/bin/bash
# round 2, GPU call I (1 GPU): mixed scenes (surfaces + lights + shadow rays) against O-gpu
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_scene.py -m gpu -q -x -s ) > gpurun_out/r02i_pytest.log 2>&1
tail -40 gpurun_out/r02i_pytest.log
