set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_r2a.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.txt
tail -30 gpurun_out/pytest_r2a.txt
python tools/l2_microbench.py --out gpurun_out/l2_microbench_r02.md > /dev/null 2>gpurun_out/l2_err.txt; tail -5 gpurun_out/l2_err.txt
python bench.py --steps 10 --warmup 3 --profile gpurun_out/kineto_n1_r2a.txt > gpurun_out/bench_n1_r2a.json 2> gpurun_out/bench_n1_r2a.err; tail -3 gpurun_out/bench_n1_r2a.err; cat gpurun_out/bench_n1_r2a.json
