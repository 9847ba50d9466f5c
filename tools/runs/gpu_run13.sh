set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/gat_repro.py 100000 256 > gpurun_out/gat_repro_hub.txt 2>&1; echo "rc=$?" >> gpurun_out/gat_repro_hub.txt
tail -4 gpurun_out/gat_repro_hub.txt | cut -c1-300
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_r2k.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2k.txt
tail -12 gpurun_out/pytest_r2k.txt
python bench.py --steps 10 --warmup 3 --profile gpurun_out/kineto_n1_r2k.txt > gpurun_out/bench_n1_r2k.json 2> gpurun_out/bench_n1_r2k.err; tail -3 gpurun_out/bench_n1_r2k.err | cut -c1-300; cut -c1-400 gpurun_out/bench_n1_r2k.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-probe --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 --profile gpurun_out/kineto_gat_yelp_n1_r2k.txt > gpurun_out/bench_gat_yelp_n1_r2k.json 2> gpurun_out/bench_gat_yelp_n1_r2k.err; tail -3 gpurun_out/bench_gat_yelp_n1_r2k.err | cut -c1-300; cut -c1-400 gpurun_out/bench_gat_yelp_n1_r2k.json
