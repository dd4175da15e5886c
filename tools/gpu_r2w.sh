#!/bin/bash
mkdir -p gpurun_out/r2w
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2w/bench_2gpu.json 2> gpurun_out/r2w/bench_2gpu.err; echo "rc=$?"
wc -lc gpurun_out/r2w/bench_2gpu.err; grep -c "NCCL INFO" gpurun_out/r2w/bench_2gpu.err; grep -E "132710400|count 33177600" gpurun_out/r2w/bench_2gpu.err | head -4 | cut -c1-200; tail -3 gpurun_out/r2w/bench_2gpu.err | cut -c1-200
python -c "
import json; j=json.load(open('gpurun_out/r2w/bench_2gpu.json')); print(j['value'], j['aux']['c5']['seconds'], j['aux']['c5']['reduce_ms'], j['aux']['c3']['value'])"
