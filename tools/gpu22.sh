set -x
timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'], d['roofline']['gemm_ms_per_step'], d['setup_s'])"
B2_NOFUSE=1 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 nofuse', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'], d['roofline']['gemm_ms_per_step'], d['setup_s'])"
timeout 200 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -3
