#!/bin/bash
mkdir -p gpurun_out
BNS_TC_TRUNC=1 timeout 240 python tools/check_dense_tc.py tn > gpurun_out/tc_tn_trunc.txt 2>&1; echo "trunc tn=$?"
BNS_TC_TRUNC=1 timeout 240 python tools/check_dense_tc.py nt > gpurun_out/tc_nt_trunc.txt 2>&1; echo "trunc nt=$?"
grep "random\|worst" gpurun_out/tc_tn_trunc.txt gpurun_out/tc_nt_trunc.txt
BNS_TC_TRUNC=1 timeout 240 python tools/check_dense_tc.py perf > gpurun_out/tc_perf_trunc.txt 2>&1; grep "tc 3xtf32\|GFLOP" gpurun_out/tc_perf_trunc.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1h.txt 2>&1; tail -n 5 gpurun_out/pytest_r1h.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_r1h_tc.json 2> gpurun_out/bench_n1_r1h_tc.err; tail -c 300 gpurun_out/bench_n1_r1h_tc.json
timeout 600 ncu --nvtx --nvtx-include "bns_timed/" --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r1h.log 2>&1; tail -n 2 gpurun_out/ncu_bench_r1h.log; wc -l gpurun_out/launches_r1h.csv
