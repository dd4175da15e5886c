#!/bin/bash
mkdir -p gpurun_out/r2g
timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py -m gpu -x -q 2>&1 | tail -2
for c in 4 8; do echo "== bounce ctas $c"; ADYPT_BOUNCE_CTAS=$c REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3"; done 2>&1 | tee gpurun_out/r2g/bounce_block_sweep.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-c4 > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2g/bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2g/bench.json'))
print('C2', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'copy_only', j['e2e']['copy_only']['value'])
print('C3', j['aux']['c3']['value'], j['aux']['c3']['roofline']['stage_ms_per_step'])
PY
