set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -3
python -m pytest tests/test_multiprocess_gpu.py -m gpu -q -k "8" > gpurun_out/pytest_mp_n8.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n8.txt
tail -12 gpurun_out/pytest_mp_n8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 40 --warmup 3 --profile gpurun_out/kineto_n8_r2g.txt > gpurun_out/bench_n8_r2g.json 2> gpurun_out/bench_n8_r2g.err; tail -3 gpurun_out/bench_n8_r2g.err | cut -c1-300; cat gpurun_out/bench_n8_r2g.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29589 bench.py --gpus 8 --steps 20 --warmup 3 --no-probe --shape ogbn-products --model gcn --n-hidden 128 --dropout 0.3 > gpurun_out/bench_products_gcn_n8_r2g.json 2> gpurun_out/bench_products_gcn_n8_r2g.err; tail -3 gpurun_out/bench_products_gcn_n8_r2g.err | cut -c1-300; cat gpurun_out/bench_products_gcn_n8_r2g.json | cut -c1-1200
