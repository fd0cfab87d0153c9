set -x
N=${1:-2}
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -5 gpurun_out/bench_c4_n$N.err; cat gpurun_out/bench_c4_n$N.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus $N --config C2 --steps 50 --warmup 3 --no-e2e 2>/dev/null | tail -1
