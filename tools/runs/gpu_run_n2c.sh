set -x
mkdir -p gpurun_out
python -m pytest tests/test_multiprocess_gpu.py -m gpu -q > gpurun_out/pytest_mp_n2c.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n2c.txt
tail -8 gpurun_out/pytest_mp_n2c.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 2 --steps 20 --warmup 3 --profile gpurun_out/kineto_n2_r2o.txt > gpurun_out/bench_n2_r2o.json 2> gpurun_out/bench_n2_r2o.err; tail -3 gpurun_out/bench_n2_r2o.err | cut -c1-300; cat gpurun_out/bench_n2_r2o.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29589 bench.py --gpus 2 --steps 10 --warmup 3 --shape yelp --model gat --n-layers 2 --n-hidden 256 --dropout 0.1 --rate 0.1 > gpurun_out/bench_gat_yelp_n2_p0.1.json 2> gpurun_out/bench_gat_yelp_n2_p0.1.err; tail -3 gpurun_out/bench_gat_yelp_n2_p0.1.err | cut -c1-300; cat gpurun_out/bench_gat_yelp_n2_p0.1.json | cut -c1-1200
