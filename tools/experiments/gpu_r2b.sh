# round 2, experiment B: sub-phase cycle profile of the dominant operators; L2 prefetch issued after the load; load split; fusion
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 900 python tools/sweep.py C4 base "pf1:B2_PF=1" "pf2:B2_PF=2" "fuse:B2_FUSE=1" "split60:B2_SPLIT=60" "split80:B2_SPLIT=80" "ns2:B2_NS=2"
timeout 600 python -m pytest tests/test_gpu_parity_large.py -q -x 2>&1 | tail -5
