#!/bin/bash
mkdir -p gpurun_out/r2d
timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py -m gpu -x -q 2>&1 | tail -3
for g in 1 2 4 8 16; do for c in 2 3 4; do echo "== group $g ctas $c"; ADYPT_PRIMARY_GROUP=$g ADYPT_PRIMARY_CTAS=$c REPS=2 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done; done 2>&1 | tee gpurun_out/r2d/primary_sweep.log
# ncu of the first-bounce shade kernel and of the bounce-0 kernel (default settings)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_bounce -s 4 -c 1 -o gpurun_out/r2d/prof_shade_bounce -f python tools/pt_time.py > gpurun_out/r2d/ncu_bounce.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_primary -s 1 -c 1 -o gpurun_out/r2d/prof_shade_primary -f python tools/pt_time.py > gpurun_out/r2d/ncu_primary.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/r2d
