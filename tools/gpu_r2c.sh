#!/bin/bash
# Round 2, GPU call C: resident kernel, fused HMC step, async host-buffer path, optimised select; full suite + bench.
O=gpurun_out/r2c; mkdir -p $O
timeout 300 python tools/sanitize_case.py > $O/fixtures_plain.log 2>&1; echo "fixtures rc=$?" >> $O/rc.txt
timeout 2400 python -m pytest tests -q -m gpu --maxfail=25 > $O/pytest_all.log 2>&1; echo "pytest_all rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 200 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -30 $O/pytest_all.log; tail -3 $O/fixtures_plain.log; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2c/bench.log').read().strip().splitlines()[-1])
    print("value", j["value"], "e2e", j["e2e"], "roofline", j["roofline"]["frac"])
    print("reduced", json.dumps(j["reduced_flop"])[:900])
    print("hmc", j["hmc"])
    print("slq", json.dumps(j["slq"])[:1500])
except Exception as e:
    print("bench parse failed", e)
PY
