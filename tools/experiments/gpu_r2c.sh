# round 2, experiment C: slab load/store microbenchmark; parity at the benchmarked configurations
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 300 tools/probe/tma_probe
timeout 900 python -m pytest tests/test_gpu_parity_large.py -q -x -s 2>&1 | grep -v "^$" | tail -40
