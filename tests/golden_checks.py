"""Checks of the library behind ``rustpde_mpi_b200`` (the CUDA build in ``-m gpu`` tests, the SIMT-emulator build in the CPU
suite) against the committed fixtures of tests/golden/ (written by tests/golden/make_golden.py from the CPU oracle).
Every function returns {name: relative max-norm error}; tolerance 1e-10 (BASELINE.json north_star)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def fixtures(prefix):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel(a, ref):
    a, ref = np.asarray(a), np.asarray(ref)
    assert a.shape == ref.shape and a.dtype == ref.dtype, (a.shape, ref.shape, a.dtype, ref.dtype)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def check_operators(b2, name):
    z = load(name)
    k0, n0, k1, n1 = (int(v) for v in z["space"])
    f = b2.Field2(b2.Space2((k0, n0), (k1, n1)))
    errs = {}
    f.v = z["forward_in"]; f.forward()
    errs["forward"] = rel(f.vhat, z["forward_out"])
    f.vhat = z["spec_in"]; f.backward()
    errs["backward"] = rel(f.v, z["backward_out"])
    f.vhat = z["spec_in"]
    errs["to_ortho"] = rel(f.to_ortho().get(), z["to_ortho_out"])
    for key in z:
        if key.startswith("gradient_"):
            d = (int(key[9]), int(key[10]))
            errs[key[:-4]] = rel(f.gradient(d, (1.5, 1.0)).get(), z[key])
    f.from_ortho(b2.DeviceArray(f.space, b2.ORTHO).set(z["from_ortho_in"]))
    errs["from_ortho"] = rel(f.vhat, z["from_ortho_out"])
    if "hholtz_adi_out" in z:
        errs["hholtz_adi"] = rel(b2.HholtzAdi(f, list(z["hholtz_adi_c"])).solve(z["solver_rhs"]).get(), z["hholtz_adi_out"])
    if "poisson_out" in z:
        # the stored host eigendecomposition (both sides of a Poisson comparison get the same one)
        eig = (z["poisson_lam"], z["poisson_fwd"], z["poisson_bwd"]) if "poisson_lam" in z else None
        x = b2.Poisson(f, [1.0, 1.0], eig=eig).solve(z["solver_rhs"]).get()
        x[0, 0] = 0
        errs["poisson"] = rel(x, z["poisson_out"])
    return errs


def check_navier(b2, name, max_steps=None):
    """``steps`` updates from the fixture's input state against the fixture's output state (max_steps: only meaningful when it
    equals the stored count -- used by the emulator case to pick fixtures it can afford)."""
    z = load(name)
    nx, ny, ra, pr, dt, aspect, periodic, steps = z["params"]
    nx, ny, steps, periodic = int(nx), int(ny), int(steps), bool(periodic)
    if max_steps is not None and steps > max_steps:
        return None
    eig = (z["poisson_lam"], z["poisson_fwd"], z["poisson_bwd"]) if "poisson_lam" in z else None
    nav = b2.Navier2D(nx, ny, float(ra), float(pr), float(dt), float(aspect), str(z["bc"]), periodic=periodic, pois_eig=eig,
                      init_random=False)
    for k in ("temp", "velx", "vely", "pres"):
        getattr(nav, k).vhat = z[f"in_{k}"]
    nav.update(steps)
    got = nav.state()
    errs = {k: rel(got[k], z[f"out_{k}"]) for k in ("temp", "velx", "vely", "pres")}
    dn = float(z["div_norm"])
    errs["div_norm"] = abs(nav.div_norm() - dn) / dn   # O(1e-3 .. 1e-1) in every fixture; a derived quantity: callers bound it by 1e-8
    nav.close()
    return errs


def check_roundtrip_and_linearity(b2, sp, seed=21):
    """Size-independent properties of the transforms on an orthonormal / Fourier space (no oracle needed, any size):
    forward(backward(c)) == c for a valid spectrum c, and forward(a u + v) == a forward(u) + forward(v)."""
    k0, n0, k1, n1 = sp
    f = b2.Field2(b2.Space2((k0, n0), (k1, n1)))
    shape, cx = f.space.shape(b2.SPECTRAL)
    rng = np.random.default_rng(seed)
    c = rng.standard_normal(shape)
    if cx:
        c = c + 1j * rng.standard_normal(shape)
        c[0] = c[0].real      # DC and Nyquist modes of a real signal are real
        c[-1] = c[-1].real
    f.vhat = c
    f.backward()
    v = f.v
    f.forward()
    errs = {"roundtrip": rel(f.vhat, c)}
    u = rng.standard_normal(v.shape)
    f.v = u
    f.forward()
    fu = f.vhat
    f.v = 0.5 * u + v
    f.forward()
    errs["linearity"] = rel(f.vhat, 0.5 * fu + c)
    f.close()
    return errs


def check_hholtz_linearity(b2, sp, seed=22):
    """HholtzAdi::solve is linear: solve(a r1 + r2) == a solve(r1) + solve(r2) (any size, no oracle needed)."""
    k0, n0, k1, n1 = sp
    f = b2.Field2(b2.Space2((k0, n0), (k1, n1)))
    hh = b2.HholtzAdi(f, [1e-3, 2e-3])
    shape, cx = f.space.shape(b2.ORTHO)
    rng = np.random.default_rng(seed)
    r1, r2 = rng.standard_normal(shape), rng.standard_normal(shape)
    if cx:
        r1 = r1 + 1j * rng.standard_normal(shape)
        r2 = r2 + 1j * rng.standard_normal(shape)
    x1, x2 = hh.solve(r1).get(), hh.solve(r2).get()
    x3 = hh.solve(0.25 * r1 + r2).get()
    hh.close(); f.close()
    return {"linearity": rel(x3, 0.25 * x1 + x2)}
