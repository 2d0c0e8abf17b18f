#!/bin/bash
# Round 2, GPU call E: ncu full captures of the resident kernel (collapsed C2, 65536 chains) and the SLQ select kernel.
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python tools/profile_case.py --case collapsed --reps 3 > $O/plain_collapsed.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 2 -o $O/resident python tools/profile_case.py --case collapsed --reps 3 > $O/ncu_collapsed.log 2>&1
echo "ncu collapsed rc=$?" >> $O/rc.txt
timeout 300 python tools/profile_case.py --case slq --reps 4 > $O/plain_slq.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slq_select -s 1 -c 2 -o $O/select python tools/profile_case.py --case slq --reps 4 > $O/ncu_slq.log 2>&1
echo "ncu slq rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/ncu_collapsed.log $O/ncu_slq.log
