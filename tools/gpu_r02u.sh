#!/bin/bash
# round 2, GPU call U (1 GPU): speed-of-light probe of a brick-staged march (TMA streaming + shared-memory sampling),
# L2 fetch granularity A/B on the texture march, K2 (dpt) at 3 / 4 CTAs per SM
mkdir -p gpurun_out
for p in brickprobe brickprobe_32x32x16 brickprobe_32x16x8; do
  echo "== $p" >> gpurun_out/r02u_brickprobe.log
  if [ $p = brickprobe ]; then timeout 120 tools/probe/$p 1024 >> gpurun_out/r02u_brickprobe.log 2>&1
  else timeout 120 tools/probe/$p 1024 quick >> gpurun_out/r02u_brickprobe.log 2>&1; fi
  echo "rc $?" >> gpurun_out/r02u_brickprobe.log
done
tail -40 gpurun_out/r02u_brickprobe.log
# DRAM bytes of one streaming pass: 3-D tensor boxes from the row-major volume (launch 2) and pre-gathered bricks (launch 9)
timeout 120 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:brickProbe -s 2 -c 1 --csv --log-file gpurun_out/r02u_probe_dram_tma3d.csv tools/probe/brickprobe 1024 quick > /dev/null 2>&1
timeout 120 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:brickProbe -s 9 -c 1 --csv --log-file gpurun_out/r02u_probe_dram_bulk1d.csv tools/probe/brickprobe 1024 quick > /dev/null 2>&1
tail -2 gpurun_out/r02u_probe_dram_tma3d.csv gpurun_out/r02u_probe_dram_bulk1d.csv
for g in 32 64 128; do
  for rate in 0.5 1.0; do
    echo "L2_FETCH=$g rate $rate: $(DVR_B200_L2_FETCH=$g timeout 120 python tools/profile_scene.py --what c2 --rate $rate --frames 40 2>&1 | tail -1 | sed 's/.*: //')" | tee -a gpurun_out/r02u_l2fetch.log
  done
done
echo "default rate 0.5: $(timeout 120 python tools/profile_scene.py --what c2 --rate 0.5 --frames 40 2>&1 | tail -1 | sed 's/.*: //')" | tee -a gpurun_out/r02u_l2fetch.log
for v in occ3 occ4; do
  echo "dpt $v: $(DVR_B200_LIB=$PWD/visrtx_b200/variants/libdvr_$v.so timeout 120 python tools/profile_scene.py --what dpt --frames 40 2>&1 | tail -1 | sed 's/.*: //')" | tee -a gpurun_out/r02u_dpt_occ.log
done
echo "dpt default: $(timeout 120 python tools/profile_scene.py --what dpt --frames 40 2>&1 | tail -1 | sed 's/.*: //')" | tee -a gpurun_out/r02u_dpt_occ.log
