#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export THY_SLQ_TIMING=1
timeout 300 $TR --nproc-per-node 2 --master-port 29701 tools/slq_probe.py --full-run --max-iterations 128000000 --warm 5 > $O/full_warm.json 2> $O/a.err
timeout 300 $TR --nproc-per-node 2 --master-port 29702 tools/slq_probe.py --rounds 2048 > $O/rounds.json 2> $O/b.err
timeout 300 $TR --nproc-per-node 2 --master-port 29703 tools/slq_probe.py --rounds 2048 --max-iterations 200000000 > $O/rounds_maxit2e8.json 2> $O/c.err
timeout 300 $TR --nproc-per-node 2 --master-port 29704 tools/slq_probe.py --full-run --max-iterations 128000000 --import-oracle > $O/full_oracle.json 2> $O/d.err
for f in a b c d; do echo "== $f"; grep -h "thy slq timing" $O/$f.err | head -3; done
cat $O/full_warm.json $O/full_oracle.json; grep -h -o '"us_per_round": [0-9.]*' $O/rounds.json $O/rounds_maxit2e8.json
