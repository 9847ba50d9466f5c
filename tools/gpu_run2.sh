set -x
mkdir -p gpurun_out
for slab in 128 64; do for chunk in 128 256 512; do
python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --slab $slab --chunk $chunk --no-cusparse --iters 10 2>/dev/null | tail -1 >> gpurun_out/spmm_slab_chunk_r2.txt
done; done
python tools/bench_spmm.py --shape reddit --parts 4 --F 256 --slab 0 --chunk 256 --no-cusparse --iters 10 2>/dev/null | tail -1 >> gpurun_out/spmm_slab_chunk_r2.txt
python tools/bench_spmm.py --shape reddit --parts 4 --F 256 --slab 64 --chunk 256 --no-cusparse --iters 10 2>/dev/null | tail -1 >> gpurun_out/spmm_slab_chunk_r2.txt
cat gpurun_out/spmm_slab_chunk_r2.txt
