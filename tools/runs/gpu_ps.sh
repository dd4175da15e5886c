#!/bin/bash
for c in 22 24; do ADYPT_PRIMARY_CTAS=$c timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -1; done
for c in 0 22 24 0 24; do echo "== primary variant $c (22 / 24 = class-sorted, 2 / 4 items per thread)"; ADYPT_PRIMARY_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/primary_sorted.log
