#!/bin/bash
O=gpurun_out/r4a; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident_dmma -s 1 -c 1 -o $O/resident64k python tools/profile_case.py --case collapsed > $O/ncu.log 2>&1
THY_SLQ_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_slq_walk$ -s 6 -c 1 -o $O/walk_final python tools/slq_probe.py --pool 125000 --k 7812 --rounds 20 > $O/ncu_walk.log 2>&1
tail -n 2 $O/ncu.log $O/ncu_walk.log; ls -la $O
