#!/bin/bash
# Builds kernel-tuning variants of the library (dev tool for the sweeps recorded in profiles/).
# usage: tools/build_variants.sh NAME "-DSMAP_FUSE_MINB=3 -DSMAP_FUSE_GATHER=4 ..." [NAME2 "..."]  -> csrc/variants/NAME.so
# knobs: SMAP_FUSE_MINB / ROUND / GATHER, SMAP_FUSE_TMA (+ STAGES, GROUP), SMAP_FUSE_PERSISTENT, SMAP_FUSE_GRID_DIV,
#        SMAP_AUX_STREAMS, SMAP_TAG_MAX_PLANES, SMAP_FUSE_STATS, SMAP_ABL_NO_{DRAIN,GATHER,SCATTER,DEFER},
#        SMAP_FUSE_PF_IMAGE, SMAP_FUSE_PF_LABEL, SMAP_FUSE_PF_CLOUD (prefetch experiments, profiles/r1k_stall_breakdown.md)
set -e
cd "$(dirname "$0")/../vision_semantic_segmentation_b200/csrc"
mkdir -p variants
while [ $# -ge 2 ]; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC \
       -Xcompiler -fvisibility=hidden -shared $2 -Xptxas -v -o variants/$1.so smap.cu 2>&1 \
       | grep -A2 "k_fuseILi1ELi1E" | grep -E "registers|spill" | sed 's/ptxas info    ://g' | tr '\n' ' '
  echo " -> variants/$1.so ($2)"
  shift 2
done
