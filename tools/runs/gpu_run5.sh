set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_r2e.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2e.txt
tail -15 gpurun_out/pytest_r2e.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile gpurun_out/kineto_n1_r2e.txt > gpurun_out/bench_n1_r2e.json 2> gpurun_out/bench_n1_r2e.err; tail -5 gpurun_out/bench_n1_r2e.err; cat gpurun_out/bench_n1_r2e.json | cut -c1-400
