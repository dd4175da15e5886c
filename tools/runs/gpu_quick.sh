#!/bin/bash
# quick confidence run after a shading change: parity tests + C3 stage times
timeout 900 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_fullsize.py tests/test_textures.py -m gpu -x -q 2>&1 | tail -2
REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"
