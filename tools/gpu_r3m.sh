#!/bin/bash
O=gpurun_out/r3n; mkdir -p $O
timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
THY_SLQ_PREDRAW=1 timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq_pre.log 2>&1; echo "pytest_slq_pre rc=$?" >> $O/rc.txt
for PD in 0 1; do
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_pre$PD.json 2>> $O/probe.err
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 250000 --k 15625 > $O/probe_n4_pre$PD.json 2>> $O/probe.err
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_pre$PD.json 2>> $O/probe.err
done
cat $O/rc.txt; tail -n 3 $O/pytest_slq.log $O/pytest_slq_pre.log; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round": [0-9.]*\|"walk_adapt": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -5 $O/probe.err
