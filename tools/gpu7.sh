set -x
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -2
python tools/opprof.py C4
B2_LN=2 python tools/opprof.py C4
B2_E=8 python tools/opprof.py C2
B2_E=8 B2_LN=2 python tools/opprof.py C2
for v in "16 4" "8 4" "8 2" "4 2" "16 2"; do set -- $v; B2_E=$1 B2_LN=$2 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 E=$1 LN=$2', d['ms_per_step'], d['roofline']['frac'])"; done
for ln in 4 2; do B2_LN=$ln python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 LN=$ln', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'])"; done
B2_LN=2 B2_E=8 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 LN=2 E=8', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'])"
for ln in 4 2; do B2_LN=$ln python bench.py --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3 LN=$ln', d['ms_per_step'], d['roofline']['frac'])"; done
