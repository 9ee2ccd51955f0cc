#!/bin/bash
# round 2, GPU call B: the whole GPU suite (no -x)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multigpu.py > gpurun_out/r02b_pytest.log 2>&1
tail -40 gpurun_out/r02b_pytest.log
