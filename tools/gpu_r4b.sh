#!/bin/bash
O=gpurun_out/r4g; mkdir -p $O
timeout 900 python -m pytest tests/test_collapse_gpu.py tests/test_parity_gpu.py tests/test_hmc_gpu.py -q -m gpu --maxfail=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python tools/bench_small.py > $O/small.json 2> $O/small.err; echo "small rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 4 $O/pytest.log
python - <<'PY'
import json
sm=json.load(open('gpurun_out/r4g/small.json'))
for k,v in sm.items():
    if isinstance(v,dict): print(k,{kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items()})
PY
