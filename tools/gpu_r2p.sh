#!/bin/bash
mkdir -p gpurun_out/r2p
timeout 600 python bench.py --steps 5 --warmup 3 --no-c4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); c=j['aux']['c3']; print('C2 clocks', j['clocks']); print('C3', c['value'], {k:round(v,2) for k,v in c['roofline']['stage_ms_per_step'].items()}, c['clocks']); print('C5', j['aux']['c5']['seconds'], j['aux']['c5']['clocks'])" | tee gpurun_out/r2p/bench_clocks.log
