#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
THY_SLQ_PREDRAW=1 timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu > $O/pytest_slq_predraw.log 2>&1; echo "pytest_slq_predraw rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu > $O/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $O/rc.txt
# single rank, N = 8 sized local work: in-place vs pre-drawn randomness
THY_SLQ_PREDRAW=0 timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_inplace.json 2>> $O/probe.err
THY_SLQ_PREDRAW=1 timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 > $O/probe_small_predraw.json 2>> $O/probe.err
THY_SLQ_PREDRAW=1 timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 > $O/probe_big_predraw.json 2>> $O/probe.err
THY_SLQ_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --only-slq > $O/only_slq_n2.json 2> $O/a.err
cat $O/rc.txt; tail -n 3 $O/pytest_slq.log $O/pytest_slq_predraw.log $O/pytest_multi.log; grep -h -o '"us_per_round": [0-9.]*\|"walk_adapt": [0-9.]*\|"pool_spread_side_stream": [0-9.]*' $O/probe_*.json; grep -h "thy slq timing" $O/a.err | head -2
python - <<'PY'
import json
s=json.loads(open("gpurun_out/r2t/only_slq_n2.json").read().strip().splitlines()[-1])["slq"]
print("N=2 bench leg", s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["within_tolerance"], s["phase_us_per_round"])
PY
