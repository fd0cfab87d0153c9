# round 2, experiment A: more warps per SM (1024-thread CTAs, 64 registers) and L2 prefetch of upcoming operands
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -3
timeout 900 python tools/sweep.py C4 base "pf1:B2_PF=1" "pf2:B2_PF=2" "pf3:B2_PF=3" "t1024:B2_T1024=1,B2_E=8" "t1024pf2:B2_T1024=1,B2_E=8,B2_PF=2" "t1024fuse:B2_T1024=1,B2_E=8,B2_PF=2,B2_FUSE=1"
timeout 100 python tools/copyprobe.py 4097
B2_T1024=1 B2_E=8 timeout 100 python tools/copyprobe.py 4097
B2_PF=2 timeout 100 python tools/copyprobe.py 4097
