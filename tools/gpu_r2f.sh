#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_hmc_gpu.py tests/test_collapse_gpu.py -q -m gpu -k "resident or hmc or collapse or fused" --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
THY_RESIDENT_WARPS=12 timeout 600 python tools/bench_small.py > $O/small_w12.json 2> $O/small_w12.err; echo "w12 rc=$?" >> $O/rc.txt
THY_RESIDENT_WARPS=8 timeout 600 python tools/bench_small.py > $O/small_w8.json 2> $O/small_w8.err; echo "w8 rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest.log; cat $O/small_w12.json; cat $O/small_w8.json | head -30
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2f/bench.log').read().strip().splitlines()[-1])
print("hmc", j["hmc"]); print("slq phases", j["slq"].get("phase_us_per_round"), j["slq"].get("seconds"))
PY
