#!/bin/bash
mkdir -p gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
for ch in 1 2 4 8; do for c in 3 4; do echo "== chunk $ch ctas $c"; ADYPT_PRIMARY_CHUNK=$ch ADYPT_PRIMARY_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done; done 2>&1 | tee gpurun_out/r2u/primary_chunk_sweep.log
