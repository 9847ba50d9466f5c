set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29588 bench.py --gpus 8 --steps 40 --warmup 3 --profile gpurun_out/kineto_n8_r2p.txt > gpurun_out/bench_n8_r2p.json 2> gpurun_out/bench_n8_r2p.err; tail -2 gpurun_out/bench_n8_r2p.err | cut -c1-300; cut -c1-600 gpurun_out/bench_n8_r2p.json
for p in 0.1 0.01 0.5 1.0; do
  extra="--no-probe"; if [ "$p" = "0.1" ]; then extra=""; fi
  timeout 120 $TR --master-port 29589 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline $extra --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 --rate $p > gpurun_out/bench_gat_yelp_n8_p$p.json 2> gpurun_out/bench_gat_yelp_n8_p$p.err; echo "rc=$?"; tail -2 gpurun_out/bench_gat_yelp_n8_p$p.err | cut -c1-300; cut -c1-300 gpurun_out/bench_gat_yelp_n8_p$p.json
done
python -m pytest tests/test_multiprocess_gpu.py -m gpu -q -k "8 and cuda-graph" > gpurun_out/pytest_mp_n8_final.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n8_final.txt
tail -5 gpurun_out/pytest_mp_n8_final.txt
timeout 100 $TR --master-port 29590 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-probe --watchdog 90 --shape papers100m --scale 0.1 --n-hidden 128 > gpurun_out/bench_papers_s0.1_n8.json 2> gpurun_out/bench_papers_s0.1_n8.err; echo "rc=$?"; tail -3 gpurun_out/bench_papers_s0.1_n8.err | cut -c1-300; cut -c1-700 gpurun_out/bench_papers_s0.1_n8.json
