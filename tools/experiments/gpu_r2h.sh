# round 2, experiment H: GEMM v2 (barrier-free pipeline, 16-k stages, greedy column split), 128-byte-row transposing stores
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -q -x 2>&1 | tail -5
timeout 600 python tools/sweep.py C4 base
SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C2 base
timeout 100 python tools/copyprobe.py 4097 0,1,8,2,3
