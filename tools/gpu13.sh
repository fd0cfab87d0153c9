B2_LN=2 python tools/copyprobe.py 4097
B2_LN=2 timeout 200 python tools/opprof.py C4
B2_LN=2 B2_NODIRECT=1 timeout 200 python tools/opprof.py C4 | head -4
B2_LN=2 timeout 200 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 LN=2', d['ms_per_step'], d['roofline']['frac'], d['roofline']['lane_ms_per_step'])"
B2_LN=2 python tools/copyprobe.py 1025
B2_LN=2 timeout 100 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 LN=2', d['ms_per_step'], d['roofline']['frac'])"
