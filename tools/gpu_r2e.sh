#!/bin/bash
mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_cli.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
for c in 4 3; do echo "== bounce ctas $c"; ADYPT_PRIMARY_GROUP=16 ADYPT_BOUNCE_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/r2e/bounce_sweep.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_bounce -s 4 -c 1 -o gpurun_out/r2e/prof_shade_bounce -f python tools/pt_time.py > gpurun_out/r2e/ncu_bounce.log 2>&1; echo "ncu1 rc=$?"
