#!/bin/bash
O=gpurun_out/r3k; mkdir -p $O
timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
for V in base A B; do
  if [ $V = base ]; then unset THYLACINE_B200_LIB; else export THYLACINE_B200_LIB=$PWD/scratch/variants/lib$V.so; fi
  timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_$V.json 2>> $O/probe.err
  timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_$V.json 2>> $O/probe.err
done
cat $O/rc.txt; tail -n 3 $O/pytest_slq.log; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round": [0-9.]*\|"walk_adapt": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -5 $O/probe.err
