set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "gat or GAT or sddmm or compact" > gpurun_out/pytest_r2m.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2m.txt
tail -12 gpurun_out/pytest_r2m.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-probe --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 --profile gpurun_out/kineto_gat_yelp_n1_r2m.txt > gpurun_out/bench_gat_yelp_n1_r2m.json 2> gpurun_out/bench_gat_yelp_n1_r2m.err; tail -3 gpurun_out/bench_gat_yelp_n1_r2m.err | cut -c1-300; cut -c1-300 gpurun_out/bench_gat_yelp_n1_r2m.json
head -40 gpurun_out/kineto_gat_yelp_n1_r2m.txt | cut -c1-200
