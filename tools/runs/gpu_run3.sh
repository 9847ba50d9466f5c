set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_r2c.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2c.txt
tail -40 gpurun_out/pytest_r2c.txt
python bench.py --steps 10 --warmup 3 --profile gpurun_out/kineto_n1_r2c.txt > gpurun_out/bench_n1_r2c.json 2> gpurun_out/bench_n1_r2c.err; tail -5 gpurun_out/bench_n1_r2c.err; cat gpurun_out/bench_n1_r2c.json
