#!/bin/bash
mkdir -p gpurun_out/r2ag
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2ag/bench_reference.json 2> gpurun_out/r2ag/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ag/bench.json 2> gpurun_out/r2ag/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ag/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ag/launches_c3.csv python tools/pt_time.py > gpurun_out/r2ag/pt_time_ncu.log 2>&1; echo "launch list c3 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
