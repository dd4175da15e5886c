#!/bin/bash
N=$1
mkdir -p gpurun_out/r2ao
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2ao/bench_${N}gpu.json 2> gpurun_out/r2ao/bench_${N}gpu.err; echo "bench$N rc=$?"
grep -E "\[bench\] C5|Bytes -> Algo" gpurun_out/r2ao/bench_${N}gpu.err | grep -E "C5|132710400" | head -3
gzip -9f gpurun_out/r2ao/bench_${N}gpu.err
head -c 300 gpurun_out/r2ao/bench_${N}gpu.json
