#!/bin/bash
mkdir -p gpurun_out/r2i
timeout 120 python tools/pool_check.py 2>&1 | tee gpurun_out/r2i/pool_check.log | tail -12; echo "check rc=${PIPESTATUS[0]}"
TUNE_VARIANTS=0,12,0,12 TUNE_THRESHOLDS=28 timeout 300 python tools/gpu_tune.py 2>&1 | tee gpurun_out/r2i/tune_pool.log | grep -E "variant|any-hit|PT "
