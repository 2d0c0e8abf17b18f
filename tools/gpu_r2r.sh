#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
THY_SLQ_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --only-slq > $O/only_slq_graph.json 2> $O/a.err
THY_SLQ_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29703 tools/slq_probe.py --full-run --max-iterations 128000000 > $O/probe_full.json 2> $O/b.err
THY_SLQ_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29704 tools/slq_probe.py --rounds 2048 > $O/probe_rounds.json 2> $O/c.err
grep -h "thy slq timing" $O/*.err; cat $O/probe_full.json; grep -h -o '"us_per_round": [0-9.]*' $O/probe_rounds.json
