#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
for S in 1 2 4; do
  THY_SLQ_WALK_TEAM=$S THY_SLQ_PREDRAW=1 THY_SLQ_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum --clock-control none -s 200 -c 80 --csv --log-file $O/launches_S$S.csv python tools/slq_probe.py --pool 125000 --k 7812 --rounds 12 > $O/ncu_S$S.log 2>&1
done
THY_SLQ_WALK_TEAM=1 THY_SLQ_PREDRAW=0 THY_SLQ_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum --clock-control none -s 200 -c 80 --csv --log-file $O/launches_inplace_S1.csv python tools/slq_probe.py --pool 125000 --k 7812 --rounds 12 > $O/ncu_inplace.log 2>&1
ls -la $O
