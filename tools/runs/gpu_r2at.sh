#!/bin/bash
mkdir -p gpurun_out/r2at
timeout 900 python tests/fullsize/c2_glsl_parity.py > gpurun_out/r2at/c2_glsl_parity.json 2> gpurun_out/r2at/c2.err; echo "c2 rc=$?"; cat gpurun_out/r2at/c2_glsl_parity.json | head -c 1200; echo
timeout 900 python tests/fullsize/c3_parity.py 64 > gpurun_out/r2at/c3_fullsize_parity.json 2> gpurun_out/r2at/c3.err; echo "c3 rc=$?"; cat gpurun_out/r2at/c3_fullsize_parity.json | head -c 1500; echo
