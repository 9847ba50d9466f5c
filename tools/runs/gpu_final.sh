#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1_final.txt 2>&1; tail -n 4 gpurun_out/pytest_r1_final.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_r1_final.json 2> gpurun_out/bench_n1_r1_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_r1_final.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','ms_per_step','eager_ms_per_step','gpu_launches']}, d['e2e']['value'], d.get('dense_roofline'), d.get('cpu_baseline'))"
