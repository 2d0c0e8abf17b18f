#!/bin/bash
O=gpurun_out/r4h; mkdir -p $O
timeout 900 python -m pytest tests/test_collapse_gpu.py tests/test_parity_gpu.py tests/test_hmc_gpu.py tests/test_abi.py tests/test_cpp_api.py -q -m gpu --maxfail=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python scratch/dbg_split.py > $O/dbg.json 2>&1
cat $O/rc.txt; tail -n 4 $O/pytest.log; cat $O/dbg.json
