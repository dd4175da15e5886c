#!/bin/bash
mkdir -p gpurun_out/r2o
timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py -m gpu -x -q 2>&1 | tail -2
for sk in 0 4096 65536 1048576; do echo "== skew $sk (pt_time)"; ADYPT_WAVEFRONT_SKEW=$sk REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/r2o/slab_skew.log
for sk in 0 65536; do echo "== skew $sk (bench)"; ADYPT_WAVEFRONT_SKEW=$sk timeout 600 python bench.py --steps 5 --warmup 3 --no-c4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); c=j['aux']['c3']; print('C3', c['value'], {k:round(v,2) for k,v in c['roofline']['stage_ms_per_step'].items()})"; done 2>&1 | tee -a gpurun_out/r2o/slab_skew.log
