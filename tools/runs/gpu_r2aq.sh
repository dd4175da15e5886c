#!/bin/bash
# slot-base flag as shared-memory atomics: timing against r2ao, then compute-sanitizer over the final kernels
mkdir -p gpurun_out/r2aq
for i in 1 2; do REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done | tee gpurun_out/r2aq/pt_time.log
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $S --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_traversal.py -m gpu -x -q -k "not c1_primary and not c2_incoherent and not chunk_schedule and not fuzz" > gpurun_out/r2aq/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2aq/memcheck.log
timeout 1200 $S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py -m gpu -x -q -k "path_tracer_matches or profiling_hooks or max_bounce" > gpurun_out/r2aq/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2aq/racecheck.log
timeout 900 $S --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_traversal.py -m gpu -x -q -k "path_tracer_matches or viewer_modes or variants_agree or golden_scenes" > gpurun_out/r2aq/initcheck.log 2>&1; echo "initcheck rc=$?"; tail -3 gpurun_out/r2aq/initcheck.log
