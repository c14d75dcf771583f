#!/bin/bash
# Dev tool (the pool has no compute-sanitizer): builds the library with -DSMAP_DEBUG_BOUNDS -- every index the kernels form
# is checked against its array and violations are counted -- runs the GPU parity tests against that build and prints what
# every test process counted.  Run ON the GPU box: gpurun -- 'bash tools/bounds_check.sh [pytest args]'
cd "$(dirname "$0")/.."
bash tools/build_variants.sh bounds "-DSMAP_DEBUG_BOUNDS" > /dev/null 2>&1 || { echo "build failed"; exit 1; }
rm -f gpurun_out/bounds_*.txt
SMAP_EXPECT_BOUNDS=1 SMAP_LIB_PATH=$PWD/vision_semantic_segmentation_b200/csrc/variants/bounds.so \
  python -m pytest ${@:-tests -m gpu -q -n 3} 2>&1 | tail -4
cat gpurun_out/bounds_*.txt
