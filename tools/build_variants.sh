#!/bin/bash
# Builds kernel-tuning variants of the library (dev tool for the sweeps recorded in profiles/).
# usage: [EXTRA="-D..."] tools/build_variants.sh "ROUND MINB" ...  -> csrc/variants/r<ROUND>_b<MINB>[_tag].so
set -e
cd "$(dirname "$0")/../vision_semantic_segmentation_b200/csrc"
mkdir -p variants
for v in "$@"; do
  set -- $v
  out=variants/r$1_b$2${TAG:+_$TAG}.so
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC \
       -Xcompiler -fvisibility=hidden -shared -DSMAP_STREAM_ROUND=$1 -DSMAP_STREAM_MINB=$2 $EXTRA \
       -Xptxas -v -o $out smap.cu 2>&1 | grep -A2 "k_streamILi0" | grep -E "registers|spill" | sed 's/ptxas info    ://g' | tr '\n' ' '
  echo " -> $out"
done
