# round 2: the default bench line with the calibrated CPU arm; the two largest configurations on one GPU (C5 = configs[4] shape,
# C6 = the north-star 8193^2 case: host LAPACK setup of two 4096^2 parity blocks, cached)
set -x
export B2_EIG_CACHE=/tmp/eig
timeout 1200 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; cut -c1-300 gpurun_out/bench_c4.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err; cut -c1-1200 gpurun_out/bench_c4_reference.json
timeout 600 python bench.py --config C5 --no-cpu-baseline --steps 10 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -2 gpurun_out/bench_c5.err; cut -c1-300 gpurun_out/bench_c5.json
timeout 1500 python bench.py --config C6 --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_c6.json 2> gpurun_out/bench_c6.err; tail -2 gpurun_out/bench_c6.err; cut -c1-300 gpurun_out/bench_c6.json
