# round 2, experiment G: own FP64 GEMM on the tiled arrays (gemm_f64.cuh) vs the library DGEMM; chunk-streaming band ops
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -q -x 2>&1 | tail -5
timeout 600 python tools/sweep.py C4 base
B2_CUBLAS=1 timeout 600 python tools/sweep.py C4 cublas
B2_NOBANDC=1 timeout 600 python tools/sweep.py C4 nobandc
SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C2 base
B2_CUBLAS=1 SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C2 cublas
