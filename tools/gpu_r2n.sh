#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
for NS in 0 2000 4000 6000 9000; do
  THY_RESIDENT_STAGGER_NS=$NS timeout 300 python - > $O/stagger_$NS.json 2>> $O/err.log <<'PY'
import json, os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_small as b
post, theta = b.problem(100, 10_000, 20260102)
col = post.collapsed()
out = {"stagger_ns": os.environ.get("THY_RESIDENT_STAGGER_NS")}
for B in (4096, 16384, 65536, 262144):
    out[str(B)] = b.time_eval(col, theta, B, steps=100)["ms_per_step"]
post1, theta1 = b.problem(10, 1000, 20260101)
out["c1_65536"] = b.time_eval(post1, theta1, 65536, steps=100)["ms_per_step"]
print(json.dumps(out))
PY
done
cat $O/stagger_*.json; tail -3 $O/err.log
