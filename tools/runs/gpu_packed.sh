#!/bin/bash
# node step with packed FFMA2 / FADD2 slab evaluations (variants 13 / 14 / 15 = 4 / 3 / 2 conversion planes on the I2F pipe), plus 128-byte nodes (16 / 17 / 18), against the product kernel
TUNE_VARIANTS=${V:-0,13,16,17,18,13,16} TUNE_THRESHOLDS=28 timeout 900 python tools/gpu_tune.py 2>&1 | grep -vE "^build|library" | tee gpurun_out/${OUT:-packed_ffma2.log}
