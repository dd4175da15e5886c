#!/bin/bash
# round 2, GPU call B (2 GPUs): gpu tests after the Sobol / counter changes, the N=2 bench line (strong legs + NCCL reduce), EGL probe 2
mkdir -p gpurun_out/r2b
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2b/pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b/bench_2gpu.json 2> gpurun_out/r2b/bench_2gpu.err; echo "bench2 rc=$?"
grep -E "\[bench\]|AllReduce.*count 33177600" gpurun_out/r2b/bench_2gpu.err | head -20
python tools/egl_probe2.py > gpurun_out/r2b/egl_probe2.txt 2>&1
nvidia-smi topo -m > gpurun_out/r2b/topo.txt 2>&1
head -c 600 gpurun_out/r2b/bench_2gpu.json
