#!/bin/bash
# Round 2, 2-GPU call: multi-GPU tests (torchrun worker + single-process threads), SLQ round-time probes, N=2 bench.
O=gpurun_out/r2l; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu > $O/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "user_link or resident" > $O/pytest_single.log 2>&1; echo "pytest_single rc=$?" >> $O/rc.txt
# probes: what a round costs with and without the graph / the collective
timeout 300 python tools/slq_probe.py --pool 125000 --k 62500 > $O/probe_w1_p125k_graph.json 2>> $O/probe.err
THY_SLQ_GRAPH=0 timeout 300 python tools/slq_probe.py --pool 125000 --k 62500 > $O/probe_w1_p125k_plain.json 2>> $O/probe.err
timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_w1_p125k_k7812.json 2>> $O/probe.err
timeout 300 $TR --nproc-per-node 2 --master-port 29702 tools/slq_probe.py > $O/probe_w2_graph.json 2>> $O/probe.err
THY_SLQ_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29703 tools/slq_probe.py > $O/probe_w2_plain.json 2>> $O/probe.err
timeout 600 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_n2.log 2> $O/bench_n2.err; echo "bench2 rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_small.py > $O/small.json 2> $O/small.err
cat $O/rc.txt; tail -4 $O/pytest_multi.log $O/pytest_single.log; grep -h us_per_round $O/probe_*.json; tail -3 $O/probe.err
