#!/bin/bash
# Regenerates profiles/spmm_traffic.json (read by bench.py for roofline.traffic) and the raw metric dumps it comes from:
# one `ncu --set full` capture of the SpMM as bench.py launches it on the headline shapes.
#   world 1: the Reddit-shape inner matrix (232,965 x 232,965, 114.6 M entries), F = 256, two source-row blocks x two
#            column slabs -> TWO kernel launches per logical SpMM (their DRAM bytes are summed)
#   world 4: a 58,242-row stand-in of one rank's inner matrix (7.4 M entries, same degree law), F = 256, one launch
# Run on ONE GPU under gpurun:   bash tools/ncu_spmm_traffic.sh
set -x
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
ncu --metrics $M --clock-control none -k regex:spmm_kernel -s 6 -c 2 --csv --log-file gpurun_out/ncu_spmm_w1_r02.csv \
    python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --col-blocks 2 --iters 2 --no-cusparse > gpurun_out/ncu_spmm_w1_r02.log 2>&1
ncu --metrics $M --clock-control none -k regex:spmm_kernel -s 3 -c 1 --csv --log-file gpurun_out/ncu_spmm_w4_r02.csv \
    python tools/bench_spmm.py --shape reddit --n 58242 --e 7435796 --parts 1 --F 256 --iters 2 --no-cusparse > gpurun_out/ncu_spmm_w4_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 6 -c 1 -o gpurun_out/prof_spmm_r02_blocked \
    python tools/bench_spmm.py --shape reddit --parts 1 --F 256 --col-blocks 2 --iters 2 --no-cusparse > /dev/null 2>&1
python tools/ncu_traffic_to_json.py
