set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for mode in 0 3 1; do python bench.py --steps 50 --warmup 3 --mode $mode --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('MODE',d['config']['schedule'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['gpu_launches'], d['clocks'])"; done
for e in 8 4; do B2_E=$e python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('E=$e', d['ms_per_step'], d['roofline']['frac'])"; done
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_c2_v2.json 2>gpurun_out/bench_c2_v2.err; cat gpurun_out/bench_c2_v2.json; tail -3 gpurun_out/bench_c2_v2.err
python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_v2.json 2>gpurun_out/bench_c4_v2.err; cat gpurun_out/bench_c4_v2.json; tail -3 gpurun_out/bench_c4_v2.err
python bench.py --config C3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_v2.json 2>gpurun_out/bench_c3_v2.err; cat gpurun_out/bench_c3_v2.json; tail -3 gpurun_out/bench_c3_v2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_v2.csv python bench.py --steps 2 --warmup 3 --mode 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 120 -c 23 -o gpurun_out/prof_c2_v2 python bench.py --steps 2 --warmup 3 --mode 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 97 -c 8 -o gpurun_out/prof_c4_v2 python bench.py --config C4 --steps 1 --warmup 3 --mode 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c4.log 2>&1
tail -3 gpurun_out/ncu_full_c4.log
ls -la gpurun_out
