#!/bin/bash
# one queue-slot reservation per block and round for the classes that always go on (ADYPT_EARLY_SLOTS, default 1) against one per 32-entry chunk (0)
timeout 900 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_fullsize.py tests/test_textures.py -m gpu -x -q 2>&1 | tail -1
for c in 1 0 1 0; do echo "== block slots $c"; ADYPT_EARLY_SLOTS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/block_slots.log
