"""Host logic on CPU: the SAME sources (lane_kernel.cuh + b200pde.cu) compiled with g++ against the
SIMT emulator of tests/emu (every CUDA thread = one OS thread) and compared with the oracle.
This covers lane-program construction, coefficient vectors, thread/chunk index algebra and the
reference's golden vectors through the C ABI -- it is NOT a product path (the package never
loads the emulator build) and says nothing about GPU results; `-m gpu` does that."""
import subprocess
import sys
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# run in a subprocess: the emulator build replaces the library for the whole process
SCRIPT = r'''
import sys
sys.path.insert(0, %r)
from tests import emu
emu.activate()
import numpy as np
import rustpde_mpi_b200 as b2
from tests import gpu_checks as g
from oracle import rustpde_oracle as o

case = sys.argv[1]
if case == "ops":
    for sp in [(1, 65, 2, 65), (4, 64, 1, 65)]:
        for fn in (g.check_roundtrip_layout, g.check_to_ortho, g.check_from_ortho, g.check_backward, g.check_forward, g.check_hholtz):
            e = fn(*sp); assert e < g.TOL, (fn.__name__, sp, e)
        for d in ((1, 0), (0, 2)):
            e = g.check_gradient(*sp, d); assert e < g.TOL, ("gradient", sp, d, e)
elif case == "poisson":
    for sp in [(2, 65, 2, 65), (4, 64, 2, 65)]:
        e = g.check_poisson(*sp); assert e < g.TOL, (sp, e)
    for sp in [(1, 65, 1, 65), (4, 64, 2, 65)]:
        e = g.check_hholtz_tensor(*sp); assert e < g.TOL, ("hholtz", sp, e)
elif case == "golden":
    # reference goldens through the C ABI: src/solver/hholtz_adi.rs:215-246 and poisson.rs:295-325
    f = b2.Field2(b2.Space2(b2.cheb_dirichlet(7), b2.cheb_dirichlet(7)))
    x = b2.HholtzAdi(f, [1.0, 1.0]).solve(np.tile(np.arange(1.0, 8.0), (7, 1))).get()
    y = np.array([[-7.083e-03, -9.025e-03, -5.210e-03, 4.146e-03, 3.520e-03],
                  [5.809e-04, 7.402e-04, 4.273e-04, -3.401e-04, -2.887e-04],
                  [1.699e-04, 2.165e-04, 1.250e-04, -9.951e-05, -8.447e-05],
                  [-1.007e-03, -1.283e-03, -7.406e-04, 5.895e-04, 5.004e-04],
                  [-6.775e-04, -8.632e-04, -4.983e-04, 3.966e-04, 3.366e-04]])
    np.testing.assert_allclose(x, y, rtol=6e-4, atol=1e-7)
    f = b2.Field2(b2.Space2(b2.cheb_dirichlet(8), b2.cheb_dirichlet(7)))
    x = b2.Poisson(f, [1.0, 1.0]).solve(np.tile(np.arange(1.0, 8.0), (8, 1))).get()
    from tests.test_oracle_golden import GOLD_P2D
    np.testing.assert_allclose(x, GOLD_P2D, atol=1.5e-6)
    # shape errors replace the reference's panics
    try:
        b2.DeviceArray(f.space, b2.ORTHO).set(np.zeros((3, 3)))
        raise SystemExit("expected a shape error")
    except b2.B2Error:
        pass
elif case == "average":
    # the reference's doc tests of average_axis / average (src/field/average.rs:12-25, 38-52) through the C ABI: Chebyshev 6 x 5,
    # v[i, j] = j  =>  average_axis(0) = [0, 1, 2, 3, 4], average() = 2; both axes against the oracle on a random field
    f = b2.Field2(b2.Space2(b2.chebyshev(6), b2.chebyshev(5)))
    f.v = np.tile(np.arange(5.0), (6, 1))
    np.testing.assert_allclose(f.average_axis(0), np.arange(5.0), rtol=0, atol=2e-15)
    assert abs(f.average() - 2.0) < 2e-15
    for sp in [(1, 65, 2, 33), (4, 64, 1, 65)]:
        fo, fg = g.mk(*sp)
        v = np.random.default_rng(3).standard_normal(fo.v.shape)
        fo.v = v; fg.v = v
        for ax in (0, 1):
            np.testing.assert_allclose(fg.average_axis(ax), o.Navier2D.average_axis(fo, ax), rtol=0, atol=1e-14)
        assert abs(fg.average() - o.Navier2D.average(fo)) < 1e-14
elif case == "div":
    # Navier2D::div (navier_eq.rs:19-24) and reset_time (navier.rs:185-187) through the host mirror
    for periodic in (False, True):
        no, ng = g.make_navier_pair(64 if periodic else 65, 65, 1e5, 1.0, 0.01, 1.0, periodic)
        no.update(); ng.update(1)
        d, dref = ng.div(), no.div()
        assert d.shape == dref.shape and d.dtype == dref.dtype
        assert float(np.abs(d - dref).max() / np.abs(dref).max()) < g.TOL
        assert abs(np.sqrt(np.sum(np.abs(d) ** 2)) - ng.div_norm()) < 1e-12 * max(1.0, ng.div_norm())
        assert ng.get_time() > 0
        ng.reset_time()
        assert ng.get_time() == 0.0
elif case == "callback":
    # integrate() + the reference's callback (navier.rs:476-480, navier_io.rs:84-147): flow files and info.txt under io_dir at the
    # save times, readable back into a fresh solver
    import tempfile, os, glob
    d = tempfile.mkdtemp()
    ng = b2.Navier2D.new_confined(65, 65, 1e5, 1.0, 0.01, 1.0, "rbc")
    ng.set_velocity(0.2, 1.0, 1.0); ng.set_temperature(0.2, 1.0, 1.0)
    ng.io_dir = os.path.join(d, "data")
    b2.integrate(ng, 0.02, 0.01)
    flows = sorted(os.path.basename(f) for f in glob.glob(os.path.join(d, "data", "flow*")))
    from rustpde_mpi_b200 import snapshot as sn
    ext = sn.default_ext()
    assert flows == ["flow00000.01" + ext, "flow00000.02" + ext], flows
    lines = open(os.path.join(d, "data", "info.txt")).read().strip().splitlines()
    assert len(lines) == 2 and abs(float(lines[1].split()[0]) - 0.02) < 1e-12 and len(lines[1].split()) == 4
    snap = sn.load_datasets(os.path.join(d, "data", "flow00000.02" + ext))   # (reading a snapshot back into a solver: the "snapshot" case)
    assert abs(float(snap["time"]) - 0.02) < 1e-12 and np.array_equal(snap["temp/vhat"], ng.temp.vhat)
    ng.write_intervall = 100.0   # navier_io.rs:98-101: only near multiples of the interval
    ng.callback()
    assert len(glob.glob(os.path.join(d, "data", "flow*"))) == 2
    assert ng.callback_from_filename(None, None, True) is None   # suppress_io
elif case == "variants":
    # the same step through the alternative data-movement paths selected by the environment of this process
    errs = g.check_navier(65, 65, 1)
    assert max(errs.values()) < g.TOL, errs
    f = b2.Field2(b2.Space2(b2.cheb_dirichlet(17), b2.fourier_r2c(16) if False else b2.cheb_dirichlet(17)))
    a = np.random.default_rng(0).standard_normal((15, 15))
    f.vhat = a
    out = np.empty((15, 15)); f.vhat_into(out)
    assert np.array_equal(out, a)
elif case == "hc":
    # bc = "hc": ChebDirichletNeumann temperature base (three-term stencil OP_STEN3, PdmaPlus2 solves OP_PDMA)
    for sp in [(2, 65, 3, 65)]:
        for fn in (g.check_to_ortho, g.check_from_ortho, g.check_forward, g.check_hholtz):
            e = fn(*sp); assert e < g.TOL, (fn.__name__, sp, e)
    errs = g.check_navier(65, 65, 1, False, bc="hc")
    assert max(errs.values()) < g.TOL, errs
    errs = g.check_navier(64, 65, 1, True, bc="hc")
    assert max(errs.values()) < g.TOL, errs
elif case == "snapshot":
    # write / read (navier_io.rs:21-62) through the C ABI: same grid = identical state, other grid = interpolate_2d + backward
    import tempfile, os
    from rustpde_mpi_b200 import snapshot as sn
    d = tempfile.mkdtemp()
    for periodic, (nx, ny), (nx2, ny2) in ((True, (64, 65), (128, 65)),):   # (the confined case runs in the GPU suite)
        a = b2.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=periodic)
        a.update(1)
        fn = os.path.join(d, f"snap{int(periodic)}.npz")
        a.write(fn)
        keys = set(sn.load_datasets(fn))
        grp = lambda g: {f"{g}/{k}" for k in (("x", "dx", "y", "dy", "v") + (("vhat_re", "vhat_im") if periodic else ("vhat",)))}
        assert keys == set().union(*(grp(g) for g in ("ux", "uy", "temp", "pres", "tempbc"))) | {"time", "ra", "pr", "nu", "ka"}, keys
        b = b2.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=periodic)
        b.read(fn)
        assert abs(b.get_time() - a.get_time()) < 1e-15
        for k, v in a.state().items():
            assert np.array_equal(b.state()[k], v), k
        assert np.abs(b.temp.v - a.temp.v).max() < 1e-12
        c = b2.Navier2D(nx2, ny2, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=periodic)
        c.read(fn)
        for k, v in a.state().items():
            want = sn.interpolate_2d(v, c.state()[k].shape, periodic)
            assert np.array_equal(c.state()[k], want), k
    # oracle restatement of interpolate_2d (src/field/io.rs:151-176)
    old = np.arange(12.0).reshape(3, 4)
    new = sn.interpolate_2d(old, (5, 3), True)
    assert new.shape == (5, 3) and np.array_equal(new[:3, :3], old[:, :3] * (4 / 2)) and not new[3:].any()
elif case == "anysize":
    # transform sizes that are not 2^k (+1): dense-matrix transforms (OP_DENSE)
    for sp in [(1, 30, 2, 23), (4, 24, 1, 19)]:
        for fn in (g.check_backward, g.check_forward, g.check_hholtz):
            e = fn(*sp); assert e < g.TOL, (fn.__name__, sp, e)
    errs = g.check_navier(27, 22, 1)
    assert max(errs.values()) < g.TOL, errs
    errs = g.check_navier(24, 19, 1, True)
    assert max(errs.values()) < g.TOL, errs
elif case == "e4":
    # 4 FFT points per thread on a 128-point lane (E = 4, TPL = 16): the layout an 8-rank run of 129 x 129 picks because of
    # its padding (pitch 160); forced here on one rank with B2_E=4.  Regression: this layout once had no compile-time-geometry
    # kernel instance and silently skipped the chunk-streaming band ops.
    for sp in [(1, 129, 2, 129)]:
        for fn in (g.check_backward, g.check_forward, g.check_to_ortho, g.check_hholtz):
            e = fn(*sp); assert e < g.TOL, (fn.__name__, sp, e)
        e = g.check_gradient(*sp, (1, 0)); assert e < g.TOL, ("gradient", sp, e)
elif case == "c2c":
    # FourierC2c on axis 0 (bases.rs:15; not on the Navier2D path): complex physical values, modes in FFT order, dense-matrix transform
    for sp in [(5, 64, 1, 33), (5, 30, 2, 23)]:
        for fn in (g.check_roundtrip_layout, g.check_forward, g.check_backward, g.check_to_ortho, g.check_from_ortho, g.check_hholtz,
                   g.check_hholtz_tensor, g.check_poisson):
            e = fn(*sp); assert e < g.TOL, (fn.__name__, sp, e)
        for d in ((1, 0), (2, 1), (3, 0)):
            e = g.check_gradient(*sp, d); assert e < g.TOL, ("gradient", sp, d, e)
    try:
        b2.Field2(b2.Space2(b2.fourier_c2c(2048), b2.cheb_dirichlet(17)))
        raise SystemExit("expected B2_ERR_UNSUPPORTED")
    except b2.B2Error:
        pass
elif case == "navier":
    errs = g.check_navier(65, 65, 1)
    assert max(errs.values()) < g.TOL, errs
    e = g.check_diagnostics(65, 65, 1)
    assert e < g.TOL, e
    errs = g.check_navier(64, 65, 1, True)
    assert max(errs.values()) < g.TOL, errs
print("ok")
''' % ROOT


@pytest.mark.parametrize("case", ["ops", "poisson", "golden", "average", "div", "callback", "navier", "hc", "snapshot", "anysize", "c2c"])
def test_emulated_host_logic(case):
    r = subprocess.run([sys.executable, "-c", SCRIPT, case], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("env", [{"B2_LDTHREADS": "1", "B2_CHW": "3"},     # combining loads on the per-thread path, 3-tile sub-chunks
                                 {"B2_NOTMA": "1", "B2_NOFAST": "1"}],    # per-thread loads/stores only, generic-geometry operators
                         ids=["ldthreads-smallchunks", "threads-generic"])
def test_emulated_path_variants(env):
    """The tuning switches select alternative implementations of the same operators; each must give the same step."""
    r = subprocess.run([sys.executable, "-c", SCRIPT, "variants"], capture_output=True, text=True, timeout=900, cwd=ROOT,
                       env=dict(os.environ, **env))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]


def test_emulated_four_points_per_thread_layout():
    r = subprocess.run([sys.executable, "-c", SCRIPT, "e4"], capture_output=True, text=True, timeout=900, cwd=ROOT,
                       env=dict(os.environ, B2_E="4"))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]
