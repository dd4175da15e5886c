#!/bin/bash
# confidence run after a shading change: parity tests + C3 stage times, three times
timeout 900 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_fullsize.py tests/test_textures.py -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done | tee gpurun_out/${OUT:-quick2.log}
