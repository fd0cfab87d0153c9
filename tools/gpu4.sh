set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for e in 16 8; do B2_E=$e python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 E=$e', d['ms_per_step'], d['roofline']['frac'])"; done
python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4', d['ms_per_step'], d['roofline'])"
python bench.py --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C3', d['ms_per_step'], d['roofline']['frac'])"
# C4: three lane kernels of one fused step (A-y pass of velx, B-x pass, C-x pass), full sections
ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 100 -c 12 -o /tmp/prof_c4 python bench.py --config C4 --steps 1 --warmup 3 --mode 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c4.log 2>&1
tail -2 gpurun_out/ncu_full_c4.log
ncu -i /tmp/prof_c4.ncu-rep --page raw --csv > gpurun_out/c4_raw.csv
ncu -i /tmp/prof_c4.ncu-rep --page details --csv > gpurun_out/c4_details.csv
ncu -i /tmp/prof_c4.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/c4_source_k1.csv 2>/dev/null
ncu -i /tmp/prof_c4.ncu-rep --page source --csv --kernel-id :::6 > gpurun_out/c4_source_k6.csv 2>/dev/null
ls -la /tmp/prof_c4.ncu-rep gpurun_out
