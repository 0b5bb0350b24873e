#!/bin/bash
# round 2, first GPU pass: new digit-set kernels, guard, probe, graphs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_precision.py -x -q -m gpu -s 2>&1 | tail -150 > gpurun_out/r2a_tests.log
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
