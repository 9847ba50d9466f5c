set -x
mkdir -p gpurun_out
python -m pytest tests/test_multiprocess_gpu.py -m gpu -q > gpurun_out/pytest_mp_n2b.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n2b.txt
tail -15 gpurun_out/pytest_mp_n2b.txt
tail -5 gpurun_out/dist_check_w2_eager_abi.log
