#!/bin/bash
O=gpurun_out/r3t; mkdir -p $O
for WP in 1 0; do for CS in 0 1; do
  THY_SLQ_WALK_PRIORITY=$WP THY_SLQ_CUB_SORT=$CS timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 --blocks 2 > $O/probe_big_wp${WP}_cub${CS}.json 2>> $O/probe.err
  THY_SLQ_WALK_PRIORITY=$WP THY_SLQ_CUB_SORT=$CS timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 --blocks 2 > $O/probe_small_wp${WP}_cub${CS}.json 2>> $O/probe.err
done; done
timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
cat $O/rc.txt; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round_blocks": [^]]*\]' $f; done; tail -5 $O/probe.err
