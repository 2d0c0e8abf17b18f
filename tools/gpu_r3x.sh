#!/bin/bash
O=gpurun_out/r3x; mkdir -p $O
timeout 300 python scratch/dbg_slq.py > $O/dbg_default.json 2> $O/err.txt
THY_SLQ_CUB_SORT=1 timeout 300 python scratch/dbg_slq.py > $O/dbg_cub.json 2>> $O/err.txt
THY_SLQ_GRAPH=0 timeout 300 python scratch/dbg_slq.py > $O/dbg_nograph.json 2>> $O/err.txt
DBG_P=400000 DBG_K=25000 timeout 300 python scratch/dbg_slq.py > $O/dbg_p400k.json 2>> $O/err.txt
DBG_P=300000 DBG_K=18750 timeout 300 python scratch/dbg_slq.py > $O/dbg_p300k.json 2>> $O/err.txt
tail -3 $O/err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3x/dbg_*.json')):
    try: j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f, j["env"])
    for r in j["rounds"]: print("   ", {k:v for k,v in r.items() if k not in ("bad_first",)})
PY
