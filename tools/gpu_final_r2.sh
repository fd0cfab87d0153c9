# round 2, final evidence on one B200: the whole GPU suite, the bench lines, the launch list of two steps and full ncu captures of the
# GEMM and of every lane pass of one step (CSV export on the box)
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; cut -c1-400 gpurun_out/bench_c4.json
timeout 300 python bench.py --config C2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json
timeout 300 python bench.py --config C3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-300 gpurun_out/bench_c3.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lane_kernel|gemm_pb" -s 120 -c 58 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
exp() { # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
}
export SWEEP_OPPROF=0 SWEEP_STEPS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pb -s 2 -c 1 -o /tmp/r02f_gemm python tools/sweep.py C4 base > gpurun_out/ncu_gemm.log 2>&1
exp r02f_gemm
timeout 1200 ncu --set full --clock-control none -k regex:lane_kernel -s 54 -c 27 -o /tmp/r02f_lane python tools/sweep.py C4 base > gpurun_out/ncu_lane.log 2>&1
exp r02f_lane
SWEEP_OPPROF=1 SWEEP_STEPS=10 timeout 300 python tools/sweep.py C4 base
B2_GEMM_DBG=1 SWEEP_OPPROF=0 SWEEP_STEPS=10 timeout 300 python tools/sweep.py C4 gemm_arithmetic_only
B2_GEMM_DBG=2 SWEEP_OPPROF=0 SWEEP_STEPS=10 timeout 300 python tools/sweep.py C4 gemm_copies_only
ls -la gpurun_out | tail -12
