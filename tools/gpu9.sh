set -x
timeout 300 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -15
timeout 200 python tools/opprof.py C4
B2_NOTMA=1 timeout 200 python tools/opprof.py C4
B2_LN=2 timeout 200 python tools/opprof.py C4
timeout 100 python tools/opprof.py C2
for v in "0 4" "0 2" "1 4"; do set -- $v; if [ $1 = 1 ]; then export B2_NOTMA=1; else unset B2_NOTMA; fi; B2_LN=$2 timeout 200 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 NOTMA=$1 LN=$2', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'])"; done
unset B2_NOTMA
for v in "16 4" "8 4" "16 2"; do set -- $v; B2_E=$1 B2_LN=$2 timeout 100 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 E=$1 LN=$2', d['ms_per_step'], d['roofline']['frac'])"; done
