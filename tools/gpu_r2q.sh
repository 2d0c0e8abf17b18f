#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
THY_SLQ_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --only-slq > $O/only_slq_graph.json 2> $O/only_slq_graph.err
THY_SLQ_TIMING=1 THY_SLQ_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 --only-slq > $O/only_slq_plain.json 2> $O/only_slq_plain.err
grep -h "thy slq timing" $O/*.err; python - <<'PY'
import json
for f in ("graph","plain"):
    s=json.loads(open(f"gpurun_out/r2q/only_slq_{f}.json").read().strip().splitlines()[-1])["slq"]
    print(f, s["seconds"], s["rounds"], 1e6*s["seconds"]/s["rounds"])
PY
