# round 2, experiment L: GEMM with register double-buffered fragments (B2_GEMM_DBG=3; 4 = arithmetic alone); smaller staging slots
# (B2_CHW) = more L1 for the coefficient / twiddle streams; new GPU tests
set -x
export B2_EIG_CACHE=/tmp/eig SWEEP_OPPROF=0
timeout 900 python tools/sweep.py C4 base "chw12:B2_CHW=12" "chw8:B2_CHW=8" "chw6:B2_CHW=6" "chw4:B2_CHW=4"
B2_GEMM_DBG=3 timeout 300 python tools/sweep.py C4 db_fragments
B2_GEMM_DBG=4 timeout 300 python tools/sweep.py C4 db_compute_only
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cpp_driver.py -q -x -k "snapshot or cpp_driver or hholtz_tensor or diagnostics" 2>&1 | tail -3
