# ncu captures behind profiles/r01 (ring-load kernel, DCT kernel, one C4 step, C2 launch list); CSV export happens on the box
# because the .ncu-rep files exceed the 64 MiB gpurun_out budget.
set -x
cd /root/repo
exp() { # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$1_source.csv.gz
}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 1 -c 1 -o /tmp/r01_ring python tools/copyprobe.py 4097 4 2 > gpurun_out/ncu_ring.log 2>&1
exp r01_ring
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 1 -c 1 -o /tmp/r01_dct python tools/copyprobe.py 4097 3 2 > gpurun_out/ncu_dct.log 2>&1
exp r01_dct
timeout 600 ncu --set full --clock-control none -k regex:lane_kernel -s 75 -c 25 -o /tmp/r01_c4step python bench.py --config C4 --steps 1 --warmup 3 --mode 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
ncu -i /tmp/r01_c4step.ncu-rep --page raw --csv > gpurun_out/r01_c4step_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 60 --csv --log-file gpurun_out/r01_launches_c2.csv python bench.py --steps 2 --warmup 3 --mode 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/ | tail -20
