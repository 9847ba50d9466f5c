set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "gat or GAT or compacted" > gpurun_out/pytest_r2h.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2h.txt
tail -25 gpurun_out/pytest_r2h.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-probe --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 > gpurun_out/bench_gat_yelp_n1_r2h.json 2> gpurun_out/bench_gat_yelp_n1_r2h.err; tail -3 gpurun_out/bench_gat_yelp_n1_r2h.err | cut -c1-300; cut -c1-600 gpurun_out/bench_gat_yelp_n1_r2h.json
BNS_GAT_FUSED=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-probe --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 > gpurun_out/bench_gat_yelp_n1_r2h_unfused.json 2> gpurun_out/bench_gat_yelp_n1_r2h_unfused.err; cut -c1-300 gpurun_out/bench_gat_yelp_n1_r2h_unfused.json
