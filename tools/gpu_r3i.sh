#!/bin/bash
# 2 GPUs: multi-GPU tests + SLQ leg of the bench at N = 2
O=gpurun_out/r4n; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu > $O/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $O/rc.txt
timeout 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --only-slq > $O/only_slq_n2.json 2> $O/a.err; echo "slq2 rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 3 $O/pytest_multi.log
python - <<'PY'
import json
s=json.loads(open("gpurun_out/r4n/only_slq_n2.json").read().strip().splitlines()[-1])["slq"]
print("N=2", s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["within_tolerance"], s["phase_us_per_round"])
PY
