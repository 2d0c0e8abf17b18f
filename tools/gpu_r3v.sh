#!/bin/bash
O=gpurun_out/r3v; mkdir -p $O
for PD in 0 1; do
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 --blocks 2 > $O/probe_small_pre$PD.json 2>> $O/probe.err
  THY_SLQ_PREDRAW=$PD timeout 300 python tools/slq_probe.py --pool 250000 --k 15625 --blocks 2 > $O/probe_n4_pre$PD.json 2>> $O/probe.err
done
for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round_blocks": [^]]*\]\|"walk_adapt": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -5 $O/probe.err
