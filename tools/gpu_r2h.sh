#!/bin/bash
# Round 2, GPU call H: full GPU suite, bench line, focused timings, ncu launch list of the bench, ncu full captures of the
# kernels that had no profile in round 1.
O=gpurun_out/r2h; mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --maxfail=25 > $O/pytest_all.log 2>&1; echo "pytest_all rc=$?" >> $O/rc.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks_bench.csv &
SMI=$!
timeout 900 python bench.py --steps 1000 --warmup 10 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
kill $SMI
timeout 600 python tools/bench_small.py > $O/small.json 2> $O/small.err; echo "small rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-slq > $O/plain_bench_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-slq > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> $O/rc.txt
prof() {  # case, kernel regex, name, skip, count
  timeout 300 python tools/profile_case.py --case $1 --reps 2 > $O/plain_$3.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c $5 -o $O/$3 python tools/profile_case.py --case $1 --reps 2 > $O/ncu_$3.log 2>&1
  echo "ncu $3 rc=$?" >> $O/rc.txt
}
prof c5_fd_literal k_fd_adjoint_prefix fd_prefix 0 1
prof c5_fd "k_fd_adjoint" fd_incremental 0 1
prof hmc_c2 "k_hmc_" hmc_elementwise 3 3
prof hmc_c1 "k_resident" hmc_c1_fused 2 2
prof lfm "k_lfm_" lfm 0 6
prof moments "k_moments" moments 0 2
prof c2 "k_prior" prior 0 1
prof collapsed4k "k_resident" resident4k 1 1
prof c3 "k_wide" wide 0 4
cat $O/rc.txt; tail -6 $O/pytest_all.log
