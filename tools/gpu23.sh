set -x
timeout 200 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -3
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
( time timeout 400 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -4 gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
