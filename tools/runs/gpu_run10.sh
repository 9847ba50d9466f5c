set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/gat_repro.py 716847 256 > gpurun_out/gat_repro.txt 2>&1; echo "rc=$?" >> gpurun_out/gat_repro.txt
tail -15 gpurun_out/gat_repro.txt
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/gat_repro.py 716847 100 > gpurun_out/gat_repro100.txt 2>&1; echo "rc=$?" >> gpurun_out/gat_repro100.txt
tail -15 gpurun_out/gat_repro100.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python tools/gat_repro.py 100000 256 > gpurun_out/gat_sanitizer.txt 2>&1; tail -40 gpurun_out/gat_sanitizer.txt
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_r2i.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2i.txt
tail -12 gpurun_out/pytest_r2i.txt
