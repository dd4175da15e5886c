#!/bin/bash
mkdir -p gpurun_out/r2aa
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_bounce -s 4 -c 1 -o gpurun_out/r2aa/prof_shade_bounce -f python tools/pt_time.py > gpurun_out/r2aa/ncu_bounce.log 2>&1; echo "ncu rc=$?"
