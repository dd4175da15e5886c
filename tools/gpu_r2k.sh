#!/bin/bash
mkdir -p gpurun_out/r2k
ADYPT_BOUNCE_CTAS=1 timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py -m gpu -x -q 2>&1 | tail -2
for c in 0 1; do echo "== bounce variant $c (1 = no regrouping)"; ADYPT_BOUNCE_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/r2k/bounce_noregroup.log
