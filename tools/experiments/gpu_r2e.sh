# round 2, re-entry baseline: full GPU tests, probe, sweep with op profile, default bench line
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -8
timeout 100 python tools/copyprobe.py 4097
timeout 600 python tools/sweep.py C4 base
SWEEP_OPPROF=0 timeout 300 python tools/sweep.py C2 base
timeout 900 python bench.py | tail -1
