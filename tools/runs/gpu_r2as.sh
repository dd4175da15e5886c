#!/bin/bash
# ncu --set full of the final shading kernels: first-bounce launch of k_shade_bounce_multi (block slot reservations) and k_shade_primary_sorted
mkdir -p gpurun_out/r2as
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_bounce -s 4 -c 1 -o gpurun_out/r2as/prof_shade_bounce -f python tools/pt_time.py > gpurun_out/r2as/ncu_bounce.log 2>&1; echo "ncu bounce rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade_primary -s 1 -c 1 -o gpurun_out/r2as/prof_shade_primary -f python tools/pt_time.py > gpurun_out/r2as/ncu_primary.log 2>&1; echo "ncu primary rc=$?"
