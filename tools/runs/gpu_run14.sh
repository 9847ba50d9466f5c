set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "gat or GAT or streaming" > gpurun_out/pytest_r2l.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2l.txt
tail -12 gpurun_out/pytest_r2l.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-probe --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 --profile gpurun_out/kineto_gat_yelp_n1_r2l.txt > gpurun_out/bench_gat_yelp_n1_r2l.json 2> gpurun_out/bench_gat_yelp_n1_r2l.err; tail -3 gpurun_out/bench_gat_yelp_n1_r2l.err | cut -c1-300; cut -c1-300 gpurun_out/bench_gat_yelp_n1_r2l.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-probe --shape papers100m --scale 0.02 --n-hidden 128 > gpurun_out/bench_papers_s002_n1.json 2> gpurun_out/bench_papers_s002_n1.err; tail -3 gpurun_out/bench_papers_s002_n1.err | cut -c1-300; cut -c1-500 gpurun_out/bench_papers_s002_n1.json
