#!/bin/bash
# round 2, GPU call T (1 GPU): profile gaps — ncu --set full of the C5 apron-brick march and of K2 (dpt), launch shares,
# Fp4/Fp8/Fp16/FpN re-timed with the batch-2 brick march
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:dvrFrameKernel -s 6 -c 2 -f -o gpurun_out/r02t_c5 python tools/profile_scene.py --what c5 --frames 10 > gpurun_out/r02t_c5_ncu.log 2>&1
tail -2 gpurun_out/r02t_c5_ncu.log
timeout 300 $NCU -k regex:dvrFrameKernel -s 6 -c 2 -f -o gpurun_out/r02t_dpt python tools/profile_scene.py --what dpt --frames 10 > gpurun_out/r02t_dpt_ncu.log 2>&1
tail -2 gpurun_out/r02t_dpt_ncu.log
for what in c5 dpt; do python tools/profile_scene.py --what $what --frames 40; done
for codec in float fp4 fp8 fp16 fpn; do python tools/profile_scene.py --what c5 --nvdb-codec $codec --frames 64; done
ls -la gpurun_out/*.ncu-rep
