#!/bin/bash
mkdir -p gpurun_out/r2s
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2s/pytest_gpu.log
