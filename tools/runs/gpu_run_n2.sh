set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests/test_multiprocess_gpu.py "tests/test_parity_gpu.py::test_training_parity_eight_partitions" "tests/test_parity_gpu.py::test_training_parity_variants" -m gpu -q > gpurun_out/pytest_mp_n2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n2.txt
tail -40 gpurun_out/pytest_mp_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --profile gpurun_out/kineto_n2_r2c.txt > gpurun_out/bench_n2_r2c.json 2> gpurun_out/bench_n2_r2c.err; tail -5 gpurun_out/bench_n2_r2c.err; cat gpurun_out/bench_n2_r2c.json
