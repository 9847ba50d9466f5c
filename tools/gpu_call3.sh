#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/check_dense_tc.py tn > gpurun_out/tc_tn.txt 2>&1; TN=$?
timeout 240 python tools/check_dense_tc.py nt > gpurun_out/tc_nt.txt 2>&1; NT=$?
echo "tn=$TN nt=$NT"
grep "structured\|random\|worst\|rror" gpurun_out/tc_tn.txt gpurun_out/tc_nt.txt | cut -c1-150
timeout 240 python tools/check_dense_tc.py perf > gpurun_out/tc_perf.txt 2>&1; grep "tc 3xtf32\|GFLOP" gpurun_out/tc_perf.txt
if [ $TN -eq 0 ] && [ $NT -eq 0 ]; then
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1j.txt 2>&1; tail -n 5 gpurun_out/pytest_r1j.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_r1j_tc.json 2> gpurun_out/bench_n1_r1j_tc.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_r1j_tc.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','ms_per_step','eager_ms_per_step','gpu_launches']}, d['e2e']['value'])"
fi
