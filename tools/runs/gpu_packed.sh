#!/bin/bash
# node step with packed FFMA2 / FADD2 slab evaluations (variants 13 / 14 / 15 = 4 / 3 / 2 conversion planes on the I2F pipe) against the product kernel
TUNE_VARIANTS=0,13,14,15,0,13 TUNE_THRESHOLDS=28 timeout 900 python tools/gpu_tune.py 2>&1 | grep -vE "^build|library" | tee gpurun_out/packed_ffma2.log
