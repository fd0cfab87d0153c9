set -x
timeout 200 python bench.py --config C2 2> gpurun_out/bench_c2_full.err | tail -1 > gpurun_out/bench_c2_full.json; python -c "import json; d=json.load(open('gpurun_out/bench_c2_full.json')); print('C2', d['ms_per_step'], d['value'], 'e2e', d['e2e'], 'cpu', d['cpu_baseline']['value'], d['roofline']['frac'])"
timeout 300 python bench.py --no-cpu-baseline 2> gpurun_out/bench_c4_e2e.err | tail -1 > gpurun_out/bench_c4_e2e.json; python -c "import json; d=json.load(open('gpurun_out/bench_c4_e2e.json')); print('C4', d['ms_per_step'], d['value'], 'e2e', d['e2e'], d['roofline']['frac'])"
