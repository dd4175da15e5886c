#!/bin/bash
O=gpurun_out/r2aw; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c3.csv python tools/pt_time.py > $O/pt_time_ncu.log 2>&1; echo "launch list c3 rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
