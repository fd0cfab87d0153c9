#!/bin/bash
# last box of round 2: the dense-transform (non-2^k) tests first, then as much of the GPU suite as the remaining budget allows
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_zz_any_size.py -q -x > gpurun_out/last_dense.log 2>&1
echo "dense exit $?" >> gpurun_out/last_dense.log
tail -5 gpurun_out/last_dense.log
timeout 200 python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/last_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/last_suite.log
tail -5 gpurun_out/last_suite.log
