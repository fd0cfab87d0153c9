"""Memory-pipeline probe: load -> [DCT] -> store passes over one n x n array; prints ms and GB/s (read + write)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustpde_mpi_b200 as b2
from rustpde_mpi_b200._lib import lib, check

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
sp = b2.Space2(b2.chebyshev(n), b2.chebyshev(n))
names = {0: "copy", 1: "copy+transpose", 4: "ring-load copy", 5: "ring-load +transpose", 8: "staged-store copy", 2: "dct copy", 3: "dct +transpose"}
modes = [int(m) for m in sys.argv[2].split(',')] if len(sys.argv) > 2 else (0, 1, 4, 5, 8, 2, 3)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for mode in modes:
    ms = C.c_double()
    check(lib().b2_debug_copy(sp._h, mode, reps, C.byref(ms)))
    byt = 2.0 * 8 * ((n + 3) // 4 * 4) ** 2
    print(f"n={n} mode={mode:2d} {names.get(mode, '?'):22s} {ms.value:8.4f} ms  {byt / ms.value / 1e6:8.1f} GB/s", flush=True)
