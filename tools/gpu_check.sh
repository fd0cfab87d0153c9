# One GPU box: parity tests, the memory-pipeline probe, per-operator cycles and the default bench line.
#   gpurun --timeout 1200 -- 'bash tools/gpu_check.sh > gpurun_out/check.log 2>&1; tail -40 gpurun_out/check.log'
set -x
timeout 300 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -3
timeout 100 python tools/copyprobe.py 4097
timeout 200 python tools/opprof.py C4
timeout 100 python bench.py --config C2 --no-cpu-baseline | tail -1
timeout 600 python bench.py | tail -1
