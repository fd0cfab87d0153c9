set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
python -m pytest tests -q -m gpu -x 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_v1.json 2> gpurun_out/bench_c2_v1.err; tail -3 gpurun_out/bench_c2_v1.err; cat gpurun_out/bench_c2_v1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c2_v1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 600 -c 4 -o gpurun_out/prof_c2_v1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
