#!/bin/bash
O=gpurun_out/r4o; mkdir -p $O
timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu -k "fused_walk" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/rc.txt
THY_SLQ_PREDRAW=1 timeout 600 python -m pytest tests/test_slq_gpu.py -q -m gpu -k "fused_walk" > $O/pytest_pre.log 2>&1; echo "pre rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 12 $O/pytest.log; tail -n 3 $O/pytest_pre.log
