#!/bin/bash
O=gpurun_out/r3d; mkdir -p $O
THY_SLQ_PREDRAW=1 THY_SLQ_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_slq_walk$ -s 6 -c 1 -o $O/walk_pre python tools/slq_probe.py --pool 125000 --k 7812 --rounds 20 > $O/ncu_pre.log 2>&1
tail -n 3 $O/ncu_pre.log; ls -la $O
