set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 --deselect tests/test_bench_shape_gpu.py > gpurun_out/pytest_r2f.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f.txt
tail -30 gpurun_out/pytest_r2f.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --shape ogbn-products --model gcn --n-hidden 128 --dropout 0.3 > gpurun_out/bench_products_gcn_n1_r2f.json 2> gpurun_out/bench_products_gcn_n1_r2f.err; tail -3 gpurun_out/bench_products_gcn_n1_r2f.err; cut -c1-700 gpurun_out/bench_products_gcn_n1_r2f.json
