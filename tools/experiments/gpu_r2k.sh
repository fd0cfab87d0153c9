# round 2, experiment K: where does the GEMM lose time? arithmetic alone / data movement alone (B2_GEMM_DBG, results invalid)
set -x
export B2_EIG_CACHE=/tmp/eig SWEEP_OPPROF=0
timeout 600 python tools/sweep.py C4 base
B2_GEMM_DBG=1 timeout 300 python tools/sweep.py C4 compute_only
B2_GEMM_DBG=2 timeout 300 python tools/sweep.py C4 copies_only
