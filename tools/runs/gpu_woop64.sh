#!/bin/bash
# experiment: 64-byte Woop rows fetched with two 256-bit loads (variant 22) against the product kernel (three 128-bit loads of 48-byte rows)
TUNE_NO_PT=1 TUNE_VARIANTS=0,22,0,22 TUNE_THRESHOLDS=28 timeout 200 python tools/gpu_tune.py 2>&1 | grep -E "variant|any-hit" | tee gpurun_out/woop64.log
