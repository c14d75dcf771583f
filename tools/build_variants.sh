#!/bin/bash
# Builds kernel-tuning variants of the library (dev tool for the sweeps recorded in profiles/).
# usage: tools/build_variants.sh "ROUND ROUNDS MINB" ...  -> vision_semantic_segmentation_b200/csrc/variants/r<ROUND>_n<ROUNDS>_b<MINB>.so
set -e
cd "$(dirname "$0")/../vision_semantic_segmentation_b200/csrc"
mkdir -p variants
for v in "$@"; do
  set -- $v
  out=variants/r$1_n$2_b$3.so
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC \
       -Xcompiler -fvisibility=hidden -shared -DSMAP_STREAM_ROUND=$1 -DSMAP_STREAM_ROUNDS=$2 -DSMAP_STREAM_MINB=$3 $EXTRA \
       -Xptxas -v -o $out smap.cu 2>&1 | grep -A2 "k_streamILi0ELi1" | grep -E "registers|spill" | tr '\n' ' '
  echo " -> $out"
done
