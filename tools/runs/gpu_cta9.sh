#!/bin/bash
# the product kernel compiled for 9 CTAs per SM (56 registers, 4 spilled words) against 8 CTAs (62 registers)
{ TUNE_NO_PT=1 TUNE_VARIANTS=0 TUNE_THRESHOLDS=28 timeout 200 python tools/gpu_tune.py 2>&1 | grep -E "variant|any-hit"
  TUNE_NO_PT=1 TUNE_CTAS=9 TUNE_VARIANTS=22 TUNE_THRESHOLDS=28,26 timeout 200 python tools/gpu_tune.py 2>&1 | grep -E "variant|any-hit" | sed 's/^/9 CTAs: /'; } | tee gpurun_out/cta9.log
