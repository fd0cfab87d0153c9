set -x
timeout 400 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -3
timeout 200 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['gemm_tflops'])"
timeout 100 python bench.py --config C2 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'], d['roofline']['gemm_ms_per_step'])"
timeout 200 python tools/opprof.py C4
