#!/bin/bash
mkdir -p gpurun_out/r2j
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_pool_kernel -s 2 -c 1 -o gpurun_out/r2j/prof_pool -f python tools/profile_variant.py 12 4 > gpurun_out/r2j/ncu_pool.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2j/ncu_pool.log
