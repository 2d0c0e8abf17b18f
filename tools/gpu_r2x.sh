#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
for S in 1 2 4; do
  THY_SLQ_WALK_TEAM=$S THY_SLQ_PREDRAW=1 timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_S$S.json 2>> $O/probe.err
  THY_SLQ_WALK_TEAM=$S THY_SLQ_PREDRAW=0 timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_inplace_S$S.json 2>> $O/probe.err
  THY_SLQ_WALK_TEAM=$S THY_SLQ_PREDRAW=1 timeout 300 python tools/slq_probe.py --pool 250000 --k 15625 > $O/probe_n4_S$S.json 2>> $O/probe.err
done
THY_SLQ_WALK_TEAM=2 THY_SLQ_PREDRAW=0 timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_inplace_S2.json 2>> $O/probe.err
THY_SLQ_WALK_TEAM=1 THY_SLQ_PREDRAW=0 timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_inplace_S1.json 2>> $O/probe.err
cat $O/rc.txt; tail -n 5 $O/pytest_slq.log; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round": [0-9.]*\|"walk_adapt": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -5 $O/probe.err
