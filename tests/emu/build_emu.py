"""Build the CPU SIMT-emulator build of the library (test infrastructure only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "rustpde_mpi_b200", "csrc", "b200pde.cu")
OUT = os.path.join(HERE, "libb200pde_emu.so")
import glob
DEPS = sorted(glob.glob(os.path.join(ROOT, "rustpde_mpi_b200", "csrc", "*"))) + [os.path.join(HERE, "cuda_emu.h"),
                                                                               os.path.join(ROOT, "include", "b200pde.h")]


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    tmp = f"{OUT}.{os.getpid()}.tmp"   # several ranks of a multi-process test may build at once: build aside, then rename
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-DB2_EMU", "-I", HERE, "-pthread", "-shared", "-fPIC",
           "-Wno-unused-function", "-o", tmp, SRC]
    subprocess.run(cmd, check=True)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
