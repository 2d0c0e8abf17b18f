#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_hmc_gpu.py tests/test_collapse_gpu.py tests/test_slq_gpu.py -q -m gpu --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python tools/bench_small.py > $O/small.json 2> $O/small.err; echo "small rc=$?" >> $O/rc.txt
timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_w1_p125k_k7812.json 2>> $O/probe.err
timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_w1_p1m.json 2>> $O/probe.err
timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 --rounds 20 > $O/plain_walk.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_slq_walk -s 6 -c 1 -o $O/walk python tools/slq_probe.py --pool 125000 --k 7812 --rounds 20 > $O/ncu_walk.log 2>&1
python tools/ncu_summary.py $O/walk.ncu-rep --out $O/ncu_summary_walk_lowocc.json >> $O/ncu_walk.log 2>&1; rm -f $O/walk.ncu-rep
cat $O/rc.txt; tail -n 4 $O/pytest.log; grep -h us_per_round $O/probe_*.json
python - <<'PY'
import json
sm=json.load(open('gpurun_out/r2m/small.json'))
for k,v in sm.items():
    if isinstance(v,dict): print(k,{kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items()})
PY
