#!/bin/bash
O=gpurun_out/r3g; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $O/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
( while true; do nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader >> $O/clocks.csv; sleep 0.5; done ) & CL=$!
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/rc.txt
kill $CL
cat $O/rc.txt; tail -n 4 $O/pytest_gpu.log; tail -n 3 $O/smoke.log; tail -c 6000 $O/bench_n1.json; tail -n 5 $O/bench_n1.err
