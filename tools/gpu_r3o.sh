#!/bin/bash
O=gpurun_out/r3o; mkdir -p $O
timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
for CS in 0 1; do
  THY_SLQ_CUB_SORT=$CS timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_cub$CS.json 2>> $O/probe.err
  THY_SLQ_CUB_SORT=$CS timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_cub$CS.json 2>> $O/probe.err
done
THY_SLQ_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $O/launches_slq_round.csv python tools/slq_probe.py --pool 1000000 --k 62500 --rounds 12 > $O/ncu_slq.log 2>&1; echo "ncu slq rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 15 $O/pytest_slq.log; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round": [0-9.]*\|"walk_adapt": [0-9.]*\|"wait_quadrature": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -5 $O/probe.err
grep -h "k_bs_" $O/launches_slq_round.csv | awk -F'","' '{print $5, $(NF)}' | sort | uniq -c | sort -rn | head -12
