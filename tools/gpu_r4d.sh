#!/bin/bash
O=gpurun_out/r4d; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resident_dmma -s 1 -c 1 -o $O/resident64k python tools/profile_case.py --case collapsed > $O/ncu.log 2>&1
tail -n 2 $O/ncu.log; ls -la $O
