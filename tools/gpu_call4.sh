#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3x_kernel -s 3 -c 1 -o gpurun_out/prof_gemm3x_tn_r1 -f python tools/check_dense_tc.py perf > gpurun_out/ncu_gemm_tn.log 2>&1; tail -n 3 gpurun_out/ncu_gemm_tn.log
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:gemm3x_kernel<\(bool\)1' -s 3 -c 1 -o gpurun_out/prof_gemm3x_nt_r1 -f python tools/check_dense_tc.py perf > gpurun_out/ncu_gemm_nt.log 2>&1; tail -n 3 gpurun_out/ncu_gemm_nt.log
timeout 600 ncu --nvtx --nvtx-include "bns_timed/" --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r1i.log 2>&1; wc -l gpurun_out/launches_r1i.csv
ls -la gpurun_out/*.ncu-rep
