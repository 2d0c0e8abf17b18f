#!/bin/bash
# Round 2, 8-GPU call: bench.py at N = 8 (weak-scaling evals/s + the SLQ leg with its phase trace) and config 3.
O=gpurun_out/r2k; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 200 --warmup 5 > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench8 rc=$?" >> $O/rc.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29808 tools/run_multi.py --only c3 --out $O/multi_c3_n8.json > $O/multi_c3_n8.log 2>&1; echo "c3 rc=$?" >> $O/rc.txt
timeout 400 $TR --nproc-per-node 4 --master-port 29604 bench.py --gpus 4 --steps 100 --warmup 5 > $O/bench_n4.log 2> $O/bench_n4.err; echo "bench4 rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -c 3000 $O/bench_n8.log; tail -5 $O/bench_n8.err; cat $O/multi_c3_n8.json
