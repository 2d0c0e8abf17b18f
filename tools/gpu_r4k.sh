#!/bin/bash
O=gpurun_out/r4k; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident_dmma -s 1 -c 1 -o $O/resident4k python tools/profile_case.py --case collapsed4k > $O/ncu4k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident_dmma -s 8 -c 1 -o $O/hmc_c1 python tools/profile_case.py --case hmc_c1 > $O/ncu_hmc.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_bs_ -s 30 -c 6 -o $O/bucket_sort python tools/slq_probe.py --pool 1000000 --k 62500 --rounds 12 > $O/ncu_bs.log 2>&1
ls -la $O; tail -n 1 $O/ncu4k.log $O/ncu_hmc.log $O/ncu_bs.log
