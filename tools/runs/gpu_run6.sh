set -x
mkdir -p gpurun_out
timeout 180 env BNS_TC_PAIR=1 python tools/check_dense_tc.py tn > gpurun_out/pair_check_tn.txt 2>&1; echo "rc=$?" >> gpurun_out/pair_check_tn.txt
tail -40 gpurun_out/pair_check_tn.txt
timeout 300 env BNS_TC_PAIR=1 python tools/check_dense_tc.py perf > gpurun_out/pair_perf.txt 2>&1; echo "rc=$?" >> gpurun_out/pair_perf.txt
timeout 300 python tools/check_dense_tc.py perf > gpurun_out/nopair_perf.txt 2>&1
grep "tc 3xtf32\|M=" gpurun_out/pair_perf.txt gpurun_out/nopair_perf.txt
