#!/bin/bash
# Round 2, GPU call A: stream-kernel fix, SLQ rebuild, full GPU suite, quick bench, sanitizer on the small fixtures.
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 300 python tools/sanitize_case.py > $O/fixtures_plain.log 2>&1; echo "fixtures_plain rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "stream" --maxfail=5 > $O/pytest_stream.log 2>&1; echo "pytest_stream rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_slq_gpu.py -q -m gpu --maxfail=8 > $O/pytest_slq.log 2>&1; echo "pytest_slq rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=15 --deselect tests/test_slq_gpu.py -k "not stream" > $O/pytest_rest.log 2>&1; echo "pytest_rest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py stream slq > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/rc.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_case.py stream slq > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -5 $O/pytest_stream.log $O/pytest_slq.log $O/pytest_rest.log
tail -3 $O/fixtures_plain.log
