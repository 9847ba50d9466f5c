set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_r2b.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.txt
tail -60 gpurun_out/pytest_r2b.txt
python bench.py --steps 10 --warmup 3 --profile gpurun_out/kineto_n1_r2b.txt > gpurun_out/bench_n1_r2b.json 2> gpurun_out/bench_n1_r2b.err; tail -5 gpurun_out/bench_n1_r2b.err; cat gpurun_out/bench_n1_r2b.json
for slab in 128 64; do for chunk in 128 256 512; do
python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --slab $slab --chunk $chunk --no-cusparse --iters 10 2>/dev/null | tail -1 >> gpurun_out/spmm_slab_chunk_r2.txt
done; done
cat gpurun_out/spmm_slab_chunk_r2.txt
for cb in 2 3; do
python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --slab 128 --chunk 256 --col-blocks $cb --no-cusparse --iters 10 2>>gpurun_out/spmm_colblocks_err.txt | tail -1 >> gpurun_out/spmm_colblocks_r2.txt
done
python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --slab 256 --chunk 256 --col-blocks 4 --no-cusparse --iters 10 2>>gpurun_out/spmm_colblocks_err.txt | tail -1 >> gpurun_out/spmm_colblocks_r2.txt
python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --slab 128 --chunk 256 --col-blocks 2 --transpose --no-cusparse --iters 10 2>>gpurun_out/spmm_colblocks_err.txt | tail -1 >> gpurun_out/spmm_colblocks_r2.txt
cat gpurun_out/spmm_colblocks_r2.txt; tail -5 gpurun_out/spmm_colblocks_err.txt
