# round 2 ncu captures: the own FP64 GEMM and every lane pass of one C4 step (CSV export on the box: the .ncu-rep files exceed the gpurun_out budget)
set -x
export B2_EIG_CACHE=/tmp/eig SWEEP_OPPROF=0 SWEEP_STEPS=1
timeout 600 python tools/sweep.py C4 base
exp() { # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$1_source.csv.gz
}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pb -s 2 -c 1 -o /tmp/r02_gemm python tools/sweep.py C4 base > gpurun_out/ncu_gemm.log 2>&1
exp r02_gemm
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 54 -c 27 -o /tmp/r02_lane python tools/sweep.py C4 base > gpurun_out/ncu_lane.log 2>&1
exp r02_lane
ls -la gpurun_out/ | tail
