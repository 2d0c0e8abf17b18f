#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x > $O/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" >> $O/rc.txt
THY_SLQ_P2P=0 timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -x > $O/pytest_multi_nccl.log 2>&1; echo "pytest_multi_nccl rc=$?" >> $O/rc.txt
THY_SLQ_TIMING=1 timeout 300 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --only-slq > $O/only_slq_p2p.json 2> $O/a.err; echo "p2p rc=$?" >> $O/rc.txt
THY_SLQ_P2P=0 timeout 300 $TR --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 --only-slq > $O/only_slq_nccl.json 2> $O/b.err; echo "nccl rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 5 $O/pytest_multi.log $O/pytest_multi_nccl.log
python - <<'PY'
import json
for f in ("p2p","nccl"):
    try:
        s=json.loads(open(f"gpurun_out/r2u/only_slq_{f}.json").read().strip().splitlines()[-1])["slq"]
        print(f, s["seconds"], 1e6*s["seconds"]/s["rounds"], s["log_evidence"], s["error"], s["within_tolerance"], s["phase_us_per_round"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 $O/a.err
