#!/bin/bash
mkdir -p gpurun_out/r2r
timeout 600 python tools/overlap_probe.py 2>&1 | tee gpurun_out/r2r/overlap_probe.log | tail -8
