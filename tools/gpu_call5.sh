#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1k.txt 2>&1; tail -n 5 gpurun_out/pytest_r1k.txt
for s in 0 64; do timeout 200 python tools/bench_spmm.py --shape reddit --parts 1 --F 44 --iters 5 --no-cusparse --slab $s 2>&1 | tail -n 1 | cut -c1-400; done | tee gpurun_out/bench_spmm_f44.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_r1k.json 2> gpurun_out/bench_n1_r1k.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_r1k.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','ms_per_step','eager_ms_per_step','gpu_launches']}, d['e2e']['value'])"
