#!/bin/bash
# round 2, GPU call A: tests, the new bench line, C3 launch list, EGL probe, box topology
mkdir -p gpurun_out/r2a
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a/pytest_gpu.log
tail -5 gpurun_out/r2a/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a/bench.err
bash tools/egl_probe.sh > gpurun_out/r2a/egl_probe.txt 2>&1
(nvidia-smi topo -m; echo; numactl -H 2>&1; echo; lscpu | head -30; echo; free -g) > gpurun_out/r2a/topology.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a/launches_c3.csv python tools/pt_time.py 64 > gpurun_out/r2a/pt_time_ncu.log 2>&1; echo "ncu rc=$?"
head -c 1500 gpurun_out/r2a/bench.json
