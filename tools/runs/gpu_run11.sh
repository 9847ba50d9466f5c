set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q > gpurun_out/pytest_r2j.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2j.txt
tail -12 gpurun_out/pytest_r2j.txt
CUDA_LAUNCH_BLOCKING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-probe --mode eager --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 > gpurun_out/bench_gat_blocking.json 2> gpurun_out/bench_gat_blocking.err; grep -v "^$" gpurun_out/bench_gat_blocking.err | tail -30 | cut -c1-250
