#!/bin/bash
O=gpurun_out/r4l; mkdir -p $O
timeout 900 python -m pytest tests/test_slq_gpu.py tests/test_full_size_gpu.py -q -m gpu --maxfail=5 -k "slq or config4" > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
timeout 300 python tools/slq_probe.py --pool 125000 --k 7812 --blocks 2 > $O/probe_small.json 2>> $O/probe.err
timeout 300 python tools/slq_probe.py --pool 1000000 --k 62500 --blocks 2 > $O/probe_big.json 2>> $O/probe.err
cat $O/rc.txt; tail -n 4 $O/pytest_slq.log; for f in $O/probe_*.json; do echo $f; grep -h -o '"us_per_round_blocks": [^]]*\]\|"radix_select_flags_values": [0-9.]*' $f | tr '\n' ' '; echo; done; tail -3 $O/probe.err
