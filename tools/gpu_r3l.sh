#!/bin/bash
# final single-GPU measurements of round 2: bench line, sustained fp64 figures, launch lists
O=gpurun_out/r3l; mkdir -p $O
( while true; do nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader >> $O/clocks.csv; sleep 0.5; done ) & CL=$!
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/rc.txt
kill $CL
timeout 600 python tools/measure_sustained.py > $O/sustained.json 2> $O/sustained.err; echo "sustained rc=$?" >> $O/rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-slq > $O/ncu_launches.log 2>&1; echo "ncu bench rc=$?" >> $O/rc.txt
THY_SLQ_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $O/launches_slq_round.csv python tools/slq_probe.py --pool 1000000 --k 62500 --rounds 12 > $O/ncu_slq.log 2>&1; echo "ncu slq rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -c 1500 $O/sustained.json; tail -n 3 $O/sustained.err
