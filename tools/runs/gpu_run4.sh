set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 --deselect tests/test_bench_shape_gpu.py > gpurun_out/pytest_r2d.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2d.txt
tail -30 gpurun_out/pytest_r2d.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile gpurun_out/kineto_n1_r2d.txt > gpurun_out/bench_n1_r2d.json 2> gpurun_out/bench_n1_r2d.err; tail -5 gpurun_out/bench_n1_r2d.err; cat gpurun_out/bench_n1_r2d.json
