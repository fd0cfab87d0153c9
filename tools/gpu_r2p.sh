# round 2: dense-matrix transforms at the reference's criterion sizes on hardware; the whole GPU suite once more
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 900 python -m pytest tests/test_gpu_zz_any_size.py -q -x 2>&1 | tail -5
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -4
