set -x
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -3 gpurun_out/bench_c4_n$N.err; python -c "import json; d=json.load(open('gpurun_out/bench_c4_n$N.json')); print('C4 n=$N', d['ms_per_step'], d['value'], d['roofline']['lane_ms_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['gemm_tflops'], d['roofline']['frac'])"
