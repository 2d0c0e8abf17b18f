#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "async" > $O/pytest_async.log 2>&1; echo "pytest_async rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 200 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_async.log; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2d/bench.log').read().strip().splitlines()[-1])
    print("value", j["value"], "e2e", j["e2e"], "roofline", j["roofline"]["frac"])
    print("reduced", json.dumps(j["reduced_flop"])[:900])
    print("hmc", j["hmc"])
    print("slq", json.dumps(j["slq"])[:1500])
except Exception as e:
    print("bench parse failed", e)
PY
