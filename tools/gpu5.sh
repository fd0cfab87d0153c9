set -x
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -3
python tools/opprof.py C2
B2_E=8 python tools/opprof.py C2
python tools/opprof.py C4
python tools/opprof.py C3
for e in 16 8 4; do B2_E=$e python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 E=$e', d['ms_per_step'], d['roofline']['frac'])"; done
python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'])"
python bench.py --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', d['ms_per_step'], d['roofline']['frac'])"
