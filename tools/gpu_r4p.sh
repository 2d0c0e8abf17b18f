#!/bin/bash
O=gpurun_out/r4p; mkdir -p $O
timeout 600 python -m pytest tests/test_collapse_gpu.py -q -m gpu > $O/pytest.log 2>&1; echo "rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 25 $O/pytest.log
