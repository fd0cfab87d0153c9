# round 2, experiment N: GEMM with a dedicated copy warp (17 warps)
set -x
export B2_EIG_CACHE=/tmp/eig SWEEP_OPPROF=0
timeout 600 python tools/sweep.py C4 copywarp
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "poisson or hholtz_tensor or navier_10" 2>&1 | tail -2
