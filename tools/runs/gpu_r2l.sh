#!/bin/bash
# round 2, final single-GPU evidence: full gpu test suite, ncu captures of the final kernels, launch lists, bench + reference arm
mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l/pytest_gpu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/r2l/prof_trace_c2 -f python tools/profile_variant.py 0 5 > gpurun_out/r2l/ncu_trace_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 3 -c 1 -o gpurun_out/r2l/prof_trace_c4 -f python tools/profile_c4.py > gpurun_out/r2l/ncu_trace_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_bounce -s 4 -c 1 -o gpurun_out/r2l/prof_shade_bounce -f python tools/pt_time.py > gpurun_out/r2l/ncu_bounce.log 2>&1; echo "ncu bounce rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_primary -s 1 -c 1 -o gpurun_out/r2l/prof_shade_primary -f python tools/pt_time.py > gpurun_out/r2l/ncu_primary.log 2>&1; echo "ncu primary rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l/launches_c3.csv python tools/pt_time.py > gpurun_out/r2l/pt_time_ncu.log 2>&1; echo "launch list c3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2l/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r2l/bench_under_ncu.json 2> gpurun_out/r2l/bench_under_ncu.err; echo "launch list bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2l/bench_reference.json 2> gpurun_out/r2l/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2l/bench.err
echo c39378b3ea64 > gpurun_out/r2l/head.txt
