"""The C++ driver over the C ABI (examples/cpp_driver/navier_rbc.cpp, the reference's examples/navier_rbc.rs without Python):
built with g++, run on the GPU, compared with the same run through the Python mirror."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    out = os.path.join(tempfile.mkdtemp(), "navier_rbc")
    libdir = os.path.join(ROOT, "rustpde_mpi_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cpp_driver", "navier_rbc.cpp"),
                    "-o", out, "-L", libdir, "-lb200pde", "-ldl", f"-Wl,-rpath,{libdir}"], check=True)
    return out


@pytest.mark.parametrize("nx,ny,steps,periodic", [(129, 129, 100, 0), (128, 65, 20, 1)])
def test_cpp_driver_matches_python_mirror(nx, ny, steps, periodic):
    import rustpde_mpi_b200 as b2
    from oracle.cpu_restated import openblas_path   # only to locate a LAPACK library for the driver's host setup

    exe = _build()
    r = subprocess.run([exe, openblas_path() or "none", str(nx), str(ny), str(steps), str(periodic)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"time=(\S+) div=(\S+) temp_sum=(\S+) temp_sumsq=(\S+)", r.stdout)
    t, div, s1, s2 = (float(x) for x in m.groups())
    nav = b2.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=bool(periodic), init_random=False)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    nav.update(steps)
    nav.temp.backward()
    v = nav.temp.v
    assert abs(t - nav.get_time()) < 1e-12
    assert abs(s2 - float((v * v).sum())) <= 1e-8 * float((v * v).sum())   # the two hosts run different LAPACK calls for the eigenbasis
    assert abs(div - nav.div_norm()) <= 1e-6 * max(1.0, nav.div_norm())
