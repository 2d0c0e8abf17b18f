#!/bin/bash
# Round 2, final 8-GPU call: bench at N = 8 (peer-memory SLQ exchange), SLQ leg alone with the NCCL fallback and at N = 4,
# config 3.
O=gpurun_out/r2v; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 200 --warmup 5 > $O/bench_n8.log 2> $O/bench_n8.err; echo "bench8 rc=$?" >> $O/rc.txt
THY_SLQ_P2P=0 timeout 300 $TR --nproc-per-node 8 --master-port 29609 bench.py --gpus 8 --only-slq > $O/only_slq_n8_nccl.json 2> $O/b.err; echo "slq8 nccl rc=$?" >> $O/rc.txt
THY_SLQ_PREDRAW=0 timeout 300 $TR --nproc-per-node 8 --master-port 29610 bench.py --gpus 8 --only-slq > $O/only_slq_n8_nopredraw.json 2> $O/c.err; echo "slq8 nopredraw rc=$?" >> $O/rc.txt
timeout 300 $TR --nproc-per-node 4 --master-port 29604 bench.py --gpus 4 --only-slq > $O/only_slq_n4.json 2> $O/d.err; echo "slq4 rc=$?" >> $O/rc.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29808 tools/run_multi.py --only c3 --out $O/multi_c3_n8.json > $O/multi_c3_n8.log 2>&1; echo "c3 rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2v/bench_n8.log').read().strip().splitlines()[-1]); s=j["slq"]
print("N=8 value", j["value"], "e2e", j["e2e"]["value"], "hmc", j["hmc"]["value"], "reduced", j["reduced_flop"]["chains_per_gpu_65536"]["value"])
print("N=8 slq", s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["within_tolerance"], s["phase_us_per_round"], s["collective"][:40])
for f in ("only_slq_n8_nccl","only_slq_n8_nopredraw","only_slq_n4"):
    try:
        s=json.loads(open(f"gpurun_out/r2v/{f}.json").read().strip().splitlines()[-1])["slq"]
        print(f, s["seconds"], 1e6*s["seconds"]/s["rounds"], s["error"], s["phase_us_per_round"])
    except Exception as e: print(f,"failed",e)
PY
