#!/bin/bash
# after dropping the implied u <= 1 test: full gpu test suite, C2 timing of the product and the scalar kernel, C3 time
mkdir -p gpurun_out/r2av
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2av/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2av/pytest_gpu.log
TUNE_VARIANTS=0,19,0 TUNE_THRESHOLDS=28 timeout 600 python tools/gpu_tune.py 2>&1 | grep -E "variant|any-hit|PT " | tee gpurun_out/r2av/tune.log
