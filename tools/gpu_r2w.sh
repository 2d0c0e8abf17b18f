#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=5 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 30 $O/pytest_slq.log
