#!/bin/bash
# 8 GPUs: SLQ leg of the bench at N = 8 on the final tree
O=gpurun_out/r4q; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29708 bench.py --gpus 8 --only-slq > $O/only_slq_n8.json 2> $O/a.err; echo "slq8 rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
s=json.loads(open("gpurun_out/r4q/only_slq_n8.json").read().strip().splitlines()[-1])["slq"]
print(8, s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["within_tolerance"], s["phase_us_per_round"])
PY
