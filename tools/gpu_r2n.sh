#!/bin/bash
mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -2
for g in 16 8 4; do for c in 3 4; do echo "== group $g ctas $c"; ADYPT_PRIMARY_GROUP=$g ADYPT_PRIMARY_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done; done 2>&1 | tee gpurun_out/r2n/primary_sweep_after_atomic_fix.log
