set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -x -q -s > gpurun_out/i8_tests.log 2>&1
tail -40 gpurun_out/i8_tests.log
