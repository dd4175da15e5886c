#!/bin/bash
mkdir -p gpurun_out/r2h
TUNE_VARIANTS=0,9,10,11,0,9 TUNE_THRESHOLDS=28 timeout 900 python tools/gpu_tune.py 2>&1 | tee gpurun_out/r2h/tune_hm_lut.log | grep -E "variant|any-hit|C3|1080|spp" 
