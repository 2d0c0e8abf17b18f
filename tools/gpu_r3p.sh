#!/bin/bash
O=gpurun_out/r3q; mkdir -p $O
for CS in 0 1; do for G in 0 1; do
  THY_SLQ_CUB_SORT=$CS THY_SLQ_GRAPH=$G timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 --blocks 3 > $O/probe_small_cub${CS}_g$G.json 2>> $O/probe.err
done; done
for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round_blocks": [^]]*\]' $f; done; tail -5 $O/probe.err
