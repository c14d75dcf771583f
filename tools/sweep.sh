#!/bin/bash
# Runs bench.py (device-resident leg only) against every library variant; prints us/frame.
cd "$(dirname "$0")/.."
for so in vision_semantic_segmentation_b200/csrc/variants/*.so; do
  for c in ${CLASSES:-5 19}; do
    out=$(SMAP_LIB_PATH=$PWD/$so python bench.py $BENCH_ARGS --classes $c --steps ${STEPS:-64} --warmup 16 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1)
    echo "$(basename $so) C=$c $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.2f us/frame  frac=%.3f kernel_ms=%.4f apply_ms=%.4f" % (d["ms_per_step"]*1e3, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["apply_kernel_ms_per_frame"]))' 2>/dev/null)"
  done
done
