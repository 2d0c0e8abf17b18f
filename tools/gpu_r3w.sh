#!/bin/bash
# final validation of the round on one GPU: all GPU tests, smoke, the bench line (+ clock record)
O=gpurun_out/r4m; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $O/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
( while true; do nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw --format=csv,noheader >> $O/clocks.csv; sleep 0.5; done ) & CL=$!
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/rc.txt
kill $CL
THY_SLQ_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $O/launches_slq_round.csv python tools/slq_probe.py --pool 1000000 --k 62500 --rounds 12 > $O/ncu_slq.log 2>&1; echo "ncu slq rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 4 $O/pytest_gpu.log; tail -n 2 $O/smoke.log; tail -n 3 $O/bench_n1.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r4m/bench_n1.json').read().strip().splitlines()[-1])
print(j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["hmc"]["value"], j["reduced_flop"]["chains_per_gpu_65536"]["value"], j["gpu_launches"], j["clocks"])
sl=j["slq"]; print(sl["seconds"], sl["error"], sl["within_tolerance"], sl["phase_us_per_round"])
PY
