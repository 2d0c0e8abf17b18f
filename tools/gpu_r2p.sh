#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29702 tools/slq_probe.py --rounds 256 --blocks 8 > $O/probe_w2_graph_blocks.json 2>> $O/probe.err
THY_SLQ_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29703 tools/slq_probe.py --rounds 256 --blocks 8 > $O/probe_w2_plain_blocks.json 2>> $O/probe.err
timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 --rounds 256 --blocks 8 > $O/probe_w1_small_blocks.json 2>> $O/probe.err
grep -h us_per_round $O/*.json; tail -3 $O/probe.err
