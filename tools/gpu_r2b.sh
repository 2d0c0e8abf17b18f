#!/bin/bash
# Round 2, GPU call B: cooperative radix select + side-stream pool spread.
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=8 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_slq.log; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2b/bench.log').read().strip().splitlines()[-1])
print(json.dumps(j.get('slq'),indent=1))
PY
