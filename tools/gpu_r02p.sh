#!/bin/bash
# round 2, GPU call P (2 GPUs): first frame of a fresh sort-last driver (every display pixel written by that launch?)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py > gpurun_out/r02p_worker.log 2>&1
grep -E "^\[|MGPU" gpurun_out/r02p_worker.log
