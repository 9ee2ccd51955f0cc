#!/bin/bash
# round 2, GPU call W (1 GPU): K2r — dpt with pixel regeneration: whole dpt suite (incl. bit-identity with the tile
# kernel), timing against the tile kernel, refill threshold and occupancy A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dpt.py -q -x ) > gpurun_out/r02w_pytest.log 2>&1
tail -6 gpurun_out/r02w_pytest.log
t() { timeout 120 python tools/profile_scene.py --what dpt --frames 40 2>&1 | tail -1 | sed 's/.*: //'; }
echo "dpt regen (default build, refill 8, 3 CTAs/SM): $(t)" | tee -a gpurun_out/r02w_dpt.log
echo "dpt tile kernel (DVR_B200_DPT_REGEN=0): $(DVR_B200_DPT_REGEN=0 t)" | tee -a gpurun_out/r02w_dpt.log
for v in refill4 refill16 refill24 refill16occ2; do
  [ -f visrtx_b200/variants/libdvr_$v.so ] && echo "dpt regen $v: $(DVR_B200_LIB=$PWD/visrtx_b200/variants/libdvr_$v.so t)" | tee -a gpurun_out/r02w_dpt.log
done
