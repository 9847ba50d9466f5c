set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests/test_multiprocess_gpu.py -m gpu -q > gpurun_out/pytest_mp_n4.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n4.txt
tail -15 gpurun_out/pytest_mp_n4.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 30 --warmup 3 --profile gpurun_out/kineto_n4_r2f.txt > gpurun_out/bench_n4_r2f.json 2> gpurun_out/bench_n4_r2f.err; tail -5 gpurun_out/bench_n4_r2f.err; cat gpurun_out/bench_n4_r2f.json
