# round 2, experiment I: GEMM v3 (register double-buffered fragments); hc / Hholtz / snapshot GPU tests
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5
timeout 600 python tools/sweep.py C4 base
SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C2 base
