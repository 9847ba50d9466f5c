#!/bin/bash
# one gpurun call: bring-up of the tcgen05 dense kernels, then (only if they are right) parity + bench with them on
mkdir -p gpurun_out
timeout 240 python tools/check_dense_tc.py tn > gpurun_out/tc_tn.txt 2>&1; TN=$?
timeout 240 python tools/check_dense_tc.py nt > gpurun_out/tc_nt.txt 2>&1; NT=$?
echo "tn=$TN nt=$NT"
grep -v "^   \|bad rows" gpurun_out/tc_tn.txt | tail -n 12; head -n 24 gpurun_out/tc_nt.txt
timeout 240 python tools/check_dense_tc.py perf > gpurun_out/tc_perf.txt 2>&1; tail -n 30 gpurun_out/tc_perf.txt
if [ $TN -eq 0 ] && [ $NT -eq 0 ]; then
  BNS_DENSE=tc timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_tc.txt 2>&1; tail -n 5 gpurun_out/pytest_tc.txt
  BNS_DENSE=tc timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_tc.json 2> gpurun_out/bench_n1_tc.err; tail -c 1500 gpurun_out/bench_n1_tc.json
fi
