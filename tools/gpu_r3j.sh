#!/bin/bash
# 8 GPUs: SLQ leg of the bench at N = 8 and N = 4
O=gpurun_out/r3u; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29708 bench.py --gpus 8 --only-slq > $O/only_slq_n8.json 2> $O/a.err; echo "slq8 rc=$?" >> $O/rc.txt
timeout 300 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --only-slq > $O/only_slq_n4.json 2> $O/b.err; echo "slq4 rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
for n in (8,4):
    s=json.loads(open(f"gpurun_out/r3u/only_slq_n{n}.json").read().strip().splitlines()[-1])["slq"]
    print(n, s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["within_tolerance"], s["phase_us_per_round"], s.get("replacements_per_round_per_rank"))
PY
