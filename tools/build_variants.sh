#!/bin/bash
# Builds kernel variants of libdvr_b200.so into visrtx_b200/variants/ for A/B timing on the GPU box:
#   tools/build_variants.sh "occ4_b4:-DDVR_OCC=4 -DDVR_BATCH=4" "occ3_b8:-DDVR_OCC=3 -DDVR_BATCH=8" ...
set -e
cd "$(dirname "$0")/../visrtx_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  d=../variants/.obj_$name; rm -rf $d; mkdir -p $d
  for f in dvr_kernels dvr_macrocell dvr_api dvr_post dvr_nvdb_bricks dvr_scene; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $f.cu -o $d/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libdvr_$name.so $d/*.o
  rm -rf $d
  echo "built variants/libdvr_$name.so ($flags)"
done
