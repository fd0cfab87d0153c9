#!/bin/bash
# First box of a follow-up session (gpurun --gpus 8 -- 'bash tools/gpu_next.sh 8'): what round 2 could only check in the SIMT emulator.
#  1. multi-GPU parity tests on N ranks (8: the E = 4 x 16-thread layout of 129-point lanes, pitch 160 / 288)
#  2. the bench line at N ranks (parity_check on the same ranks, collective allocator's offset check on hardware)
#  3. (any N) the fixtures of tests/golden through the CUDA path, FourierC2c on hardware, the bench line's `ops` on slabs (--ops-multi)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w_golden_fixtures.py tests/test_gpu_zz_any_size.py -q -m gpu > gpurun_out/next_golden_c2c.log 2>&1; tail -3 gpurun_out/next_golden_c2c.log
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/next_multi_n$N.log 2>&1; tail -3 gpurun_out/next_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 20 --warmup 5 --ops-multi \
  > gpurun_out/next_bench_c4_n$N.json 2> gpurun_out/next_bench_c4_n$N.err; tail -3 gpurun_out/next_bench_c4_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/next_bench_c4_n$N.json").read().strip().splitlines()[-1])
    print(d.get("ops_error"), (d.get("ops") or {}).get("ms_per_transform"), (d.get("ops") or {}).get("ms_per_solve"))
    print({k: d[k] for k in ("n_gpus", "ms_per_step", "value", "parity_check")}, d["roofline"]["lane_ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["clocks"])
except Exception as e:
    print("no bench line:", e)
PY
