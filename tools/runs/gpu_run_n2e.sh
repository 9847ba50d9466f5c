set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29590 bench.py --gpus 2 --steps 5 --warmup 3 --shape papers100m --scale 0.01 --n-hidden 128 --watchdog 100 --no-cpu-baseline > gpurun_out/bench_papers_dbg_n2.json 2> gpurun_out/bench_papers_dbg_n2.err; echo "rc=$?"; grep -v "site-packages\|\^\^" gpurun_out/bench_papers_dbg_n2.err | tail -60 | cut -c1-200; cut -c1-900 gpurun_out/bench_papers_dbg_n2.json
