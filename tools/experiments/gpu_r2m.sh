# round 2, experiment M: GEMM warp-skew bound (B2_GEMM_GATE = 1, 2, 3)
set -x
export B2_EIG_CACHE=/tmp/eig SWEEP_OPPROF=0
timeout 600 python tools/sweep.py C4 gate2
B2_GEMM_GATE=1 timeout 300 python tools/sweep.py C4 gate1
B2_GEMM_GATE=3 timeout 300 python tools/sweep.py C4 gate3
