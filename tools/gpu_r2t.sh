#!/bin/bash
mkdir -p gpurun_out/r2t
ADYPT_BOUNCE_CTAS=16 timeout 300 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py -m gpu -x -q 2>&1 | tail -2
for c in 0 16; do echo "== bounce variant $c (16 = 64-thread blocks)"; ADYPT_BOUNCE_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/r2t/bounce_block64.log
