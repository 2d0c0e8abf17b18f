#!/bin/bash
O=gpurun_out/r3a; mkdir -p $O
export THY_SLQ_PREDRAW=1 THY_SLQ_GRAPH=0
THY_SLQ_WALK_SPLIT=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slq_walk_split -s 6 -c 1 -o $O/split4 python tools/slq_probe.py --pool 125000 --k 7812 --rounds 20 > $O/ncu_split4.log 2>&1
THY_SLQ_WALK_SPLIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slq_walk\< -s 6 -c 1 -o $O/old python tools/slq_probe.py --pool 125000 --k 7812 --rounds 20 > $O/ncu_old.log 2>&1
ls -la $O; tail -3 $O/ncu_split4.log $O/ncu_old.log
