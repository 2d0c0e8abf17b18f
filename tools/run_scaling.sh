#!/usr/bin/env bash
# The multi-GPU measurements behind profiles/r01_scaling_1_2_4_8.json and profiles/r01_config4_slq_full_1_2_4_8.json,
# on one 8-GPU box:   bash tools/run_scaling.sh [outdir]
#   bench.py (BASELINE config 2, weak scaling) at N = 1, 2, 4, 8
#   tools/run_multi.py c4full (config 4 at full size: 10^6 live points to termination) at N = 1, 2, 4, 8
#   tools/run_multi.py c3 (config 3: 65 536 pool members as 8 sub-pools) at N = 8
# THY_SLQ_TRACE=1 (device phases, graph replay off) or =2 (host phases) adds a per-round SLQ breakdown on stderr.
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29600 + N)) bench.py --gpus $N --steps 200 --warmup 10 > "$OUT/bench_n$N.log" 2>&1
  timeout 400 $TR --nproc-per-node $N --master-port $((29700 + N)) tools/run_multi.py --only c4full --out "$OUT/multi_c4full_n$N.json" > "$OUT/multi_c4full_n$N.log" 2>&1
done
timeout 400 $TR --nproc-per-node 8 --master-port 29808 tools/run_multi.py --only c3 --out "$OUT/multi_c3_n8.json" > "$OUT/multi_c3_n8.log" 2>&1
python bench.py --gpus 1 --steps 200 --warmup 10 > "$OUT/bench_n1.log" 2>&1
python tools/run_multi.py --only c4full --out "$OUT/multi_c4full_n1.json" > "$OUT/multi_c4full_n1.log" 2>&1
tail -c 400 "$OUT"/bench_n*.log
