#!/bin/bash
# 8 GPUs: the full bench line at N = 8 (what the driver's scaling run will launch)
O=gpurun_out/r4j; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 200 --warmup 5 > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench8 rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r4j/bench_n8.log').read().strip().splitlines()[-1]); s=j["slq"]
print("N=8 value", j["value"], "e2e", j["e2e"]["value"], "hmc", j["hmc"]["value"], "reduced 64k", j["reduced_flop"]["chains_per_gpu_65536"]["value"], "reduced 4k", j["reduced_flop"]["chains_per_gpu_4096"]["value"])
print("N=8 slq", s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["within_tolerance"], s["phase_us_per_round"])
PY
tail -3 $O/bench_n8.err
