#!/bin/bash
# compute-sanitizer over the final kernels of round 2 (packed / 96-byte-node traversal, class-sorted bounce 0, block slot reservations)
mkdir -p gpurun_out/r2ap
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $S --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_russian_roulette.py tests/test_gpu_traversal.py -m gpu -x -q -k "not c1_primary and not c2_incoherent and not chunk_schedule and not fuzz" > gpurun_out/r2ap/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2ap/memcheck.log
timeout 1200 $S --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py -m gpu -x -q -k "path_tracer_matches or profiling_hooks or max_bounce" > gpurun_out/r2ap/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2ap/racecheck.log; grep -c "hazard" gpurun_out/r2ap/racecheck.log
timeout 900 $S --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_traversal.py -m gpu -x -q -k "path_tracer_matches or viewer_modes or variants_agree or golden_scenes" > gpurun_out/r2ap/initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/r2ap/initcheck.log
