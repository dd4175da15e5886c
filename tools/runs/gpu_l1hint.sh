#!/bin/bash
# L1 policy hints on the Woop / node loads and the redundant u <= 1 test, as experimental builds (tools/build_variant.py) against the product library
for lib in "" woop_noalloc woop_evictfirst node_evictlast noalloc_evictlast drop_ule1 ""; do
  echo "== ${lib:-product}"
  if [ -n "$lib" ]; then export TUNE_LIB=adypt_b200/lib/variants/$lib/libadypt_b200.so; else unset TUNE_LIB; fi
  TUNE_NO_PT=1 TUNE_VARIANTS=0 TUNE_THRESHOLDS=28 timeout 300 python tools/gpu_tune.py 2>&1 | grep -E "variant|any-hit"
done | tee gpurun_out/l1hint.log
