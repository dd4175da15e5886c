#!/bin/bash
# round 2, GPU call F (8 GPUs): SoA queues sanity + N=8 bench line + PCIe ceiling with 8 ranks + NCCL log of the C5 reduce + render-group test on >1 device
mkdir -p gpurun_out/r2f
timeout 600 python -m pytest tests/test_gpu_tracer.py tests/test_gpu_glsl_golden.py tests/test_gpu_traversal.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -v -k "render_group" 2>&1 | tee gpurun_out/r2f/pytest_render_group_8gpu.log | tail -4
REPS=3 timeout 300 python tools/pt_time.py 2>&1 | grep -E "stage|C3" | tee gpurun_out/r2f/pt_time.log
nvidia-smi topo -m > gpurun_out/r2f/topo.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" >> gpurun_out/r2f/topo.txt
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/pcie_probe_multi.py 2>/dev/null | tail -1 | tee -a gpurun_out/r2f/pcie_probe_multi.jsonl
done
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=ALL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2f/bench_8gpu.json 2> gpurun_out/r2f/bench_8gpu.err; echo "bench8 rc=$?"
grep -E "\[bench\]" gpurun_out/r2f/bench_8gpu.err | tail -5
grep -E "NVLS|AllReduce|Connected all|nChannels|comm 0x.*rank 0 .*Init COMPLETE" gpurun_out/r2f/bench_8gpu.err | grep -E " \[0\] | 0 \[" | head -60 > gpurun_out/r2f/nccl_rank0_extract.log
wc -l gpurun_out/r2f/bench_8gpu.err gpurun_out/r2f/nccl_rank0_extract.log
gzip -9 gpurun_out/r2f/bench_8gpu.err
head -c 400 gpurun_out/r2f/bench_8gpu.json
