#!/bin/bash
O=gpurun_out/r3y; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 6 $O/pytest_gpu.log
