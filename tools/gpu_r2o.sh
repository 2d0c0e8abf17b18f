#!/bin/bash
# Round 2, second 8-GPU call: bench at N = 8 with the load-balance statistic, config 3 from a warm-started pool.
O=gpurun_out/r2o; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 200 --warmup 5 > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench8 rc=$?" >> $O/rc.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29808 tools/run_multi.py --only c3 --out $O/multi_c3_n8.json > $O/multi_c3_n8.log 2>&1; echo "c3 rc=$?" >> $O/rc.txt
kill $SMI
cat $O/rc.txt; cat $O/multi_c3_n8.json
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2o/bench_n8.log').read().strip().splitlines()[-1]); s=j["slq"]
print("value", j["value"], "slq seconds", s["seconds"], "us/round", 1e6*s["seconds"]/s["rounds"], s["replacements_per_round_per_rank"], s["phase_us_per_round"])
PY
