#!/bin/bash
mkdir -p gpurun_out/r2x
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $S --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py -m gpu -x -q > gpurun_out/r2x/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2x/memcheck.log
timeout 1200 $S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py -m gpu -x -q -k "path_tracer_matches or profiling_hooks or max_bounce" > gpurun_out/r2x/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2x/racecheck.log
timeout 900 $S --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py -m gpu -x -q -k "path_tracer_matches or viewer_modes" > gpurun_out/r2x/initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/r2x/initcheck.log
ADYPT_EXPERIMENTAL=1 timeout 600 $S --tool racecheck --error-exitcode 7 python tools/pool_check.py > gpurun_out/r2x/racecheck_pool.log 2>&1; echo "racecheck pool rc=$?"; tail -4 gpurun_out/r2x/racecheck_pool.log
