# round 2, experiment D: per-warp copy pipelines (stores + combining loads), stencil-on-load, scale-at-store
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_parity_large.py 2>&1 | tail -5
timeout 900 python tools/sweep.py C4 base "ldthreads:B2_LDTHREADS=1" "noldsten:B2_NOLDSTEN=1" "chw16:B2_CHW=16" "chw8:B2_CHW=8"
timeout 100 python tools/copyprobe.py 4097 0,1,4,5,8,2,3
SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C2 base "ldthreads:B2_LDTHREADS=1"
SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C3 base "ldthreads:B2_LDTHREADS=1"
timeout 900 python -m pytest tests/test_gpu_parity_large.py -q -x -s 2>&1 | grep -v "^$" | tail -15
