#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2c/pytest_gpu.log
timeout 600 python tools/pt_time.py 2>&1 | tee gpurun_out/r2c/pt_time.log
