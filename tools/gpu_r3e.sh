#!/bin/bash
O=gpurun_out/r3e; mkdir -p $O
timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
for PD in 0 1; do
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_pre$PD.json 2>> $O/probe.err
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_pre$PD.json 2>> $O/probe.err
done
THY_SLQ_PREDRAW=1 THY_SLQ_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum --clock-control none -k regex:k_slq_walk -s 20 -c 4 --csv --log-file $O/launches_small_pre1.csv python tools/slq_probe.py --pool 125000 --k 7812 --rounds 12 > $O/ncu1.log 2>&1
cat $O/rc.txt; tail -n 5 $O/pytest_slq.log; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round": [0-9.]*\|"walk_adapt": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -5 $O/probe.err
grep -h "k_slq_walk" $O/launches_small_pre1.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | head
