#!/bin/bash
mkdir -p gpurun_out/r2aj
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2aj/bench_reference.json 2> gpurun_out/r2aj/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aj/bench.json 2> gpurun_out/r2aj/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2aj/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aj/launches_c3.csv python tools/pt_time.py > gpurun_out/r2aj/pt_time_ncu.log 2>&1; echo "launch list c3 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
