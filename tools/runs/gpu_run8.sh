set -x
mkdir -p gpurun_out
bash tools/ncu_spmm_traffic.sh > gpurun_out/ncu_traffic.log 2>&1
tail -20 gpurun_out/ncu_traffic.log
for slab in 0 128 64; do
python tools/bench_spmm.py --shape reddit --n 58242 --e 7435796 --parts 1 --F 256 --slab $slab --no-cusparse --iters 20 2>/dev/null | tail -1 >> gpurun_out/spmm_n4shape_r2.txt
done
python tools/bench_spmm.py --shape reddit --n 58242 --e 7435796 --parts 1 --F 256 --slab 0 --chunk 128 --no-cusparse --iters 20 2>/dev/null | tail -1 >> gpurun_out/spmm_n4shape_r2.txt
python tools/bench_spmm.py --shape reddit --n 58242 --e 7435796 --parts 1 --F 256 --slab 128 --chunk 128 --no-cusparse --iters 20 2>/dev/null | tail -1 >> gpurun_out/spmm_n4shape_r2.txt
python tools/bench_spmm.py --shape reddit --n 29121 --e 1800000 --parts 1 --F 256 --slab 0 --no-cusparse --iters 20 2>/dev/null | tail -1 >> gpurun_out/spmm_n4shape_r2.txt
python tools/bench_spmm.py --shape reddit --n 29121 --e 1800000 --parts 1 --F 256 --slab 128 --chunk 128 --no-cusparse --iters 20 2>/dev/null | tail -1 >> gpurun_out/spmm_n4shape_r2.txt
cat gpurun_out/spmm_n4shape_r2.txt
