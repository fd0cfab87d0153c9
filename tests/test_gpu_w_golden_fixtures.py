"""(Named test_gpu_w_* so that under `pytest -x -m gpu` it runs after the older GPU suites and before the any-size file.)
The committed fixtures of tests/golden/ (inputs + oracle outputs frozen by tests/golden/make_golden.py):
 * CPU: the oracle that is checked out still reproduces them (drift pin), and the library's host logic reproduces them through
   the C ABI in the SIMT emulator (a subset it can afford);
 * GPU (-m gpu): the CUDA path through the C ABI reproduces every one of them to 1e-10."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import golden_checks as gc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPERATOR_FIXTURES = gc.fixtures("operators_")
NAVIER_FIXTURES = gc.fixtures("navier_")


def test_fixture_set_is_complete():
    from tests.golden import make_golden as mk

    assert OPERATOR_FIXTURES == sorted("operators_" + mk.space_name(sp) for sp in mk.OPERATOR_SPACES)
    assert NAVIER_FIXTURES == sorted(mk.NAVIER_CASES)


@pytest.mark.parametrize("name", OPERATOR_FIXTURES)
def test_oracle_reproduces_operator_fixture(name):
    from tests.golden import make_golden as mk

    z = gc.load(name)
    now = mk.operators(tuple(int(v) for v in z["space"]))
    assert sorted(now) == sorted(z)
    for k, v in now.items():
        if k.startswith("poisson_") and k != "poisson_out":
            continue   # the stored eigendecomposition is an input of the Navier fixtures; LAPACK builds may order / scale it differently
        assert gc.rel(np.asarray(v), z[k]) < (1e-9 if k == "poisson_out" else 1e-13), k


@pytest.mark.parametrize("name", NAVIER_FIXTURES)
def test_oracle_reproduces_navier_fixture(name):
    """From the stored input state and the stored eigendecomposition: independent of this host's LAPACK."""
    from oracle import rustpde_oracle as o

    z = gc.load(name)
    nx, ny, ra, pr, dt, aspect, periodic, steps = z["params"]
    eig = (z["poisson_lam"], z["poisson_fwd"], z["poisson_bwd"]) if "poisson_lam" in z else None
    nav = o.Navier2D(int(nx), int(ny), float(ra), float(pr), float(dt), float(aspect), str(z["bc"]), periodic=bool(periodic), pois_eig=eig)
    for k in ("temp", "velx", "vely", "pres"):
        getattr(nav, k).vhat = z[f"in_{k}"].copy()
    for _ in range(int(steps)):
        nav.update()
    for k, v in nav.state().items():
        assert gc.rel(v, z[f"out_{k}"]) < 1e-12, k
    assert abs(nav.div_norm() - float(z["div_norm"])) < 1e-12 * float(z["div_norm"])


EMU_SCRIPT = r'''
import sys
sys.path.insert(0, %r)
from tests import emu
emu.activate()
import rustpde_mpi_b200 as b2
from tests import golden_checks as gc
for name in sys.argv[1:]:
    e = gc.check_operators(b2, name) if name.startswith("operators_") else gc.check_navier(b2, name)
    dn = e.pop("div_norm", 0.0)
    assert max(e.values()) < gc.TOL and dn < 1e-8, (name, e, dn)
    print("ok", name, max(e.values()))
'''


def test_emulated_library_reproduces_fixtures():
    """Host logic (lane programs, coefficient vectors, index algebra) of the same sources, compiled for the SIMT emulator, against
    the fixtures -- test infrastructure, says nothing about GPU results."""
    from tests.emu import build_emu

    build_emu.build()
    names = ["operators_cd65_cd65", "operators_r2c64_cn65", "operators_cn65_cdn65", "operators_cd9_cd9", "operators_cn17_cd17", "operators_r2c16_cd17",
             "operators_r2c32_cn33", "navier_confined_rbc_65_random"]
    r = subprocess.run([sys.executable, "-c", EMU_SCRIPT % ROOT, *names], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok ") == len(names)


@pytest.mark.gpu
@pytest.mark.parametrize("name", OPERATOR_FIXTURES)
def test_gpu_operators_against_fixture(name):
    import rustpde_mpi_b200 as b2

    errs = gc.check_operators(b2, name)
    assert max(errs.values()) < gc.TOL, errs


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAVIER_FIXTURES)
def test_gpu_navier_against_fixture(name):
    import rustpde_mpi_b200 as b2

    errs = gc.check_navier(b2, name)
    dn = errs.pop("div_norm")
    assert max(errs.values()) < gc.TOL and dn < 1e-8, (errs, dn)


@pytest.mark.gpu
def test_gpu_average_doctest_goldens():
    """The reference's doc tests of ``average_axis`` / ``average`` (src/field/average.rs:12-25, 38-52) on the device reductions:
    Chebyshev 6 x 5, v[i, j] = j  =>  average_axis(0) = [0, 1, 2, 3, 4], average() = 2; both axes against the oracle."""
    import rustpde_mpi_b200 as b2
    from oracle import rustpde_oracle as o
    from tests import gpu_checks as g

    f = b2.Field2(b2.Space2(b2.chebyshev(6), b2.chebyshev(5)))
    f.v = np.tile(np.arange(5.0), (6, 1))
    np.testing.assert_allclose(f.average_axis(0), np.arange(5.0), rtol=0, atol=1e-14)
    assert abs(f.average() - 2.0) < 1e-14
    for sp in [(1, 65, 2, 129), (4, 64, 1, 65), (2, 1025, 1, 257)]:
        fo, fg = g.mk(*sp)
        v = np.random.default_rng(3).standard_normal(fo.v.shape)
        fo.v = v; fg.v = v
        for ax in (0, 1):
            np.testing.assert_allclose(fg.average_axis(ax), o.Navier2D.average_axis(fo, ax), rtol=0, atol=1e-13)
        assert abs(fg.average() - o.Navier2D.average(fo)) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [False, True])
def test_gpu_div_reset_time_and_reference_callback(periodic, tmp_path):
    """``Navier2D::div`` (navier_eq.rs:19-24) against the oracle, ``reset_time`` (navier.rs:185-187) and ``integrate`` with the
    reference's callback (navier.rs:476-480, navier_io.rs:84-147): flow files at the save times + info.txt."""
    import glob

    import rustpde_mpi_b200 as b2
    from rustpde_mpi_b200 import snapshot as sn
    from tests import gpu_checks as g

    no, ng = g.make_navier_pair(128 if periodic else 129, 129, 1e5, 1.0, 0.01, 1.0, periodic)
    ng.io_dir = str(tmp_path / "data")
    b2.integrate(ng, 0.03, 0.01)
    for _ in range(3):
        no.update()
    d, dref = ng.div(), no.div()
    assert d.shape == dref.shape and d.dtype == dref.dtype
    assert float(np.abs(d - dref).max() / np.abs(dref).max()) < 1e-9   # a derivative of the 1e-10 state
    assert abs(np.sqrt(np.sum(np.abs(d) ** 2)) - ng.div_norm()) < 1e-12 * max(1.0, ng.div_norm())
    ext = sn.default_ext()
    flows = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "data" / "flow*")))
    assert flows == [f"flow00000.0{k}{ext}" for k in (1, 2, 3)], flows
    lines = open(tmp_path / "data" / "info.txt").read().strip().splitlines()
    assert len(lines) == 3
    t, nu, nuv, re = (float(v) for v in lines[-1].split())
    assert abs(t - 0.03) < 1e-12
    for got, ref in ((nu, no.eval_nu()), (nuv, no.eval_nuvol()), (re, no.eval_re())):
        assert abs(got - ref) <= 1e-9 * abs(ref)
    ng.reset_time()
    assert ng.get_time() == 0.0


FULL = {"C4-size": (0, 4097, 0, 4097), "C5-size": (4, 8192, 0, 4097), "C2-size": (0, 1025, 0, 1025), "C3-size": (4, 2048, 0, 1025)}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FULL))
def test_gpu_full_size_roundtrip_and_linearity(name):
    """BASELINE.json's full sizes through size-independent properties (no oracle at these sizes): forward(backward(c)) == c on the
    orthonormal / Fourier spaces of the configurations, and the forward transform is linear."""
    import rustpde_mpi_b200 as b2

    errs = gc.check_roundtrip_and_linearity(b2, FULL[name])
    assert max(errs.values()) < gc.TOL, errs


@pytest.mark.gpu
@pytest.mark.parametrize("sp", [(1, 4097, 1, 4097), (4, 2048, 1, 1025)], ids=["C4-size", "C3-size"])
def test_gpu_full_size_hholtz_linearity(sp):
    import rustpde_mpi_b200 as b2

    errs = gc.check_hholtz_linearity(b2, sp)
    assert max(errs.values()) < gc.TOL, errs
