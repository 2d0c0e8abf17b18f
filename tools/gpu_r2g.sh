#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_hmc_gpu.py tests/test_collapse_gpu.py tests/test_full_size_gpu.py -q -m gpu --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python tools/bench_small.py > $O/small.json 2> $O/small.err; echo "small rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-slq > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest.log
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2g/small.json'))
for k,v in j.items():
    if isinstance(v,dict): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items()})
j=json.loads(open('gpurun_out/r2g/bench.log').read().strip().splitlines()[-1])
print("value", j["value"], "roofline", j["roofline"]["frac"], j["roofline"]["kernel_ms"], "e2e", j["e2e"]["value"]); print("hmc", j["hmc"]["value"])
PY
