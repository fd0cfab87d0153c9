set -x
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -2
python tools/opprof.py C4
B2_LN=2 python tools/opprof.py C4
B2_E=16 B2_LN=2 python tools/opprof.py C4
python tools/opprof.py C2
B2_E=16 python tools/opprof.py C2
for v in "8 4" "8 2" "16 2"; do set -- $v; B2_E=$1 B2_LN=$2 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 E=$1 LN=$2', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'])"; done
for v in "8 4" "8 2" "4 4"; do set -- $v; B2_E=$1 B2_LN=$2 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 E=$1 LN=$2', d['ms_per_step'], d['roofline']['frac'])"; done
