# round 2, experiment J: GEMM v4 (bounded warp skew); C++ driver over the C ABI
set -x
export B2_EIG_CACHE=/tmp/eig
SWEEP_OPPROF=0 timeout 600 python tools/sweep.py C4 base
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "poisson or hholtz_tensor or navier_10" 2>&1 | tail -3
g++ -O2 -std=c++17 -I include examples/cpp_driver/navier_rbc.cpp -o /tmp/navier_rbc -L rustpde_mpi_b200 -lb200pde -ldl -Wl,-rpath,$PWD/rustpde_mpi_b200
timeout 300 /tmp/navier_rbc $(python -c "from oracle.cpu_restated import openblas_path; print(openblas_path())") 129 129 100 0
timeout 300 /tmp/navier_rbc none 128 65 20 1
