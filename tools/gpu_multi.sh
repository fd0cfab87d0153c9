# N GPUs of one box: multi-GPU parity tests, then the default bench line (C4) at N.  If the tensor-map peer stores fault on this
# box, the per-thread peer stores (B2_PEER_THREADS=1) are measured instead and the log says so.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi.sh 2 > gpurun_out/multi2.log 2>&1; tail -30 gpurun_out/multi2.log'
set -x
N=${1:-2}
export B2_EIG_CACHE=/tmp/eig
if timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tee gpurun_out/multi_tests_n$N.log | tail -3 | grep -q " passed"; then
  echo "PEER STORES: tensor maps (default)"
else
  grep -E "pass failed|B2Error" gpurun_out/multi_tests_n$N.log | head -4 | cut -c1-400
  export B2_PEER_THREADS=1
  echo "PEER STORES: falling back to per-thread stores (B2_PEER_THREADS=1)"
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -5 gpurun_out/bench_c4_n$N.err | cut -c1-300
python -c "import json; d=json.load(open('gpurun_out/bench_c4_n$N.json')); r=d['roofline']; print('C4 n=$N ms/step', d['ms_per_step'], 'steps/s', d['value'], 'lane ms', r['lane_ms_per_step'], 'gemm ms', r['gemm_ms_per_step'], 'gemm TF', r['gemm']['achieved'], 'frac', r['frac'], 'parity', d['parity_check'])"
