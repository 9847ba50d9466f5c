set -x
mkdir -p gpurun_out
for s in 0.01 0.05; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29590 bench.py --gpus 2 --steps 5 --warmup 3 --shape papers100m --scale $s --n-hidden 128 > gpurun_out/bench_papers_s${s}_n2.json 2> gpurun_out/bench_papers_s${s}_n2.err; echo "rc=$?"; tail -4 gpurun_out/bench_papers_s${s}_n2.err | cut -c1-400; cut -c1-900 gpurun_out/bench_papers_s${s}_n2.json
nvidia-smi --query-gpu=memory.used --format=csv | head -3
done
