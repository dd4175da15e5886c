#!/bin/bash
# shared-memory part of the traversal stack: 8 entries per lane (product, 15 KB per CTA) against 6 / 5 / 4 (13 / 12 / 11 KB: at 4 the SM's
# shared-memory carve-out drops a step and the L1 grows)
for lib in "" stack6 stack5 stack4 ""; do
  echo "== ${lib:-product}"
  if [ -n "$lib" ]; then export TUNE_LIB=adypt_b200/lib/variants/$lib/libadypt_b200.so; else unset TUNE_LIB; fi
  TUNE_NO_PT=1 TUNE_VARIANTS=0 TUNE_THRESHOLDS=28 timeout 200 python tools/gpu_tune.py 2>&1 | grep -E "variant|any-hit"
done | tee gpurun_out/smem_stack.log
