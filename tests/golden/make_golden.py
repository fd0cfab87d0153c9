"""Writes the golden fixtures of tests/golden/ (run from the repo root: ``python tests/golden/make_golden.py``).

What they are: seeded inputs and the outputs of the CPU oracle (``oracle/rustpde_oracle.py``, the numpy restatement of
the reference's hot path) for the field operators, the solvers and whole ``Navier2D::update()`` steps
(/root/reference/src/navier_stokes/navier.rs:438-466), frozen as ``.npz`` files.  The reference itself is Rust and cannot be
built or imported in this image (no cargo / rustc, funspace not vendored), so these vectors are NOT outputs of the
reference: they pin the oracle against drift (a change to the oracle that moves any of them fails
``tests/test_gpu_w_golden_fixtures.py`` on the CPU) and give the GPU tests a target that does not depend on the oracle code that
happens to be checked out.  The vectors the reference's own tests hold (hholtz_adi.rs:193-246, poisson.rs:275-361,
fdma_tensor.rs:386-401, pdma_plus2.rs:210) are in tests/test_oracle_golden.py and pin the oracle itself.

The eigendecomposition of the Poisson operator handed to both sides (Poisson parity) is stored in
the fixture, so a different LAPACK build does not move the vectors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import rustpde_oracle as o  # noqa: E402

# (kind, n) pairs: ChebDirichlet 1, ChebNeumann 2, Chebyshev 0, ChebDirichletNeumann 3, FourierR2c 4
# every base pairing the reference uses (cd x cd, cn x cd, cn x cn, ch x ch, r2c x {cd, cn, ch}; cn x cdn for bc = "hc"),
# at transform sizes (65, 129) and at small sizes (9 .. 33, the dense-matrix transform path)
OPERATOR_SPACES = [(1, 65, 1, 65), (2, 65, 2, 65), (0, 65, 0, 65), (4, 64, 1, 65), (4, 64, 2, 65), (2, 65, 3, 65), (1, 129, 2, 65),
                   (2, 65, 1, 65), (4, 64, 0, 65), (1, 9, 1, 9), (2, 17, 1, 17), (4, 16, 1, 17), (0, 33, 0, 33), (4, 32, 2, 33)]
GRADIENTS = [(1, 0), (0, 1), (2, 0), (0, 2), (1, 1)]
NAVIER_CASES = {
    # name: (nx, ny, ra, pr, dt, aspect, bc, periodic, init, steps)
    "navier_confined_rbc_65": (65, 65, 1e5, 1.0, 0.01, 1.0, "rbc", False, "modes", 5),
    "navier_confined_rbc_65_random": (65, 65, 1e4, 1.0, 0.01, 1.0, "rbc", False, "random", 2),
    "navier_confined_hc_65": (65, 65, 1e5, 1.0, 0.01, 1.0, "hc", False, "modes", 3),
    "navier_periodic_rbc_64x65": (64, 65, 1e5, 1.0, 0.01, 1.0, "rbc", True, "modes", 5),
    "navier_confined_rbc_129x65_aspect2": (129, 65, 1e6, 0.7, 0.005, 2.0, "rbc", False, "modes", 3),
    # BASELINE configs[0] = examples/navier_rbc.rs: 129 x 129, Ra 1e5, Pr 1, dt 0.01, set_velocity / set_temperature(0.2, 1, 1), 100 steps
    "navier_c1_129_100steps": (129, 129, 1e5, 1.0, 0.01, 1.0, "rbc", False, "modes", 100),
}


def space_name(sp):
    names = {0: "ch", 1: "cd", 2: "cn", 3: "cdn", 4: "r2c"}
    return f"{names[sp[0]]}{sp[1]}_{names[sp[2]]}{sp[3]}"


def rand_like(shape, dtype, rng):
    a = rng.standard_normal(shape)
    if dtype == np.complex128:
        a = a + 1j * rng.standard_normal(shape)
    return a


def spec_input(fo, seed):
    """Random composite coefficients with a physically sized decay; DC / Nyquist rows of an r2c axis real."""
    rng = np.random.default_rng(seed)
    a = rand_like(fo.vhat.shape, fo.vhat.dtype, rng)
    if fo.vhat.dtype == np.complex128:
        a[0] = a[0].real
        a[-1] = a[-1].real
    i = np.arange(a.shape[0])[:, None]
    j = np.arange(a.shape[1])[None, :]
    return a / (1.0 + i + j) ** 2


def poisson_eig_host(kind0, n0, c0):
    """Host eigendecomposition for the confined Poisson solve: the same routine the package's host mirror uses (it only needs
    the shared library for the two small operator matrices, no GPU)."""
    import rustpde_mpi_b200 as b2

    return b2.poisson_eig(kind0, n0, c0)


def operators(sp):
    k0, n0, k1, n1 = sp
    out = {"space": np.array(sp)}
    fo = o.Field2(o.Space2(o.Base(k0, n0), o.Base(k1, n1)))
    rng = np.random.default_rng(11)
    v = rng.uniform(-0.1, 0.1, fo.v.shape)
    fo.v = v.copy(); fo.forward()
    out["forward_in"], out["forward_out"] = v, fo.vhat.copy()
    a = spec_input(fo, 12)
    fo.vhat = a.copy(); fo.backward()
    out["spec_in"], out["backward_out"] = a, fo.v.copy()
    fo.vhat = a.copy()
    out["to_ortho_out"] = np.array(fo.to_ortho())
    osh = out["to_ortho_out"].shape
    b = rand_like(osh, fo.vhat.dtype, np.random.default_rng(13))
    fo.from_ortho(b.copy())
    out["from_ortho_in"], out["from_ortho_out"] = b, fo.vhat.copy()
    fo.vhat = a.copy()
    for d in GRADIENTS:
        out[f"gradient_{d[0]}{d[1]}_out"] = np.array(fo.gradient(d, (1.5, 1.0)))
    if k0 != 0 and k1 != 0:   # solvers are defined on composite (Galerkin) bases
        rhs = rand_like(osh, fo.vhat.dtype, np.random.default_rng(14))
        out["solver_rhs"] = rhs
        out["hholtz_adi_c"] = np.array([0.02, 0.03])
        out["hholtz_adi_out"] = np.array(o.HholtzAdi(fo, [0.02, 0.03]).solve(rhs.copy()))
        if k0 in (1, 2, 4) and k1 in (1, 2):
            eig = poisson_eig_host(k0, n0, 1.0) if k0 in (1, 2) else None
            if eig is not None:
                out["poisson_lam"], out["poisson_fwd"], out["poisson_bwd"] = eig
            x = np.array(o.Poisson(fo, [1.0, 1.0], eig=eig).solve(rhs.copy()))
            x[0, 0] = 0   # the shifted-singular mode is removed by the caller (navier_eq.rs:161)
            out["poisson_out"] = x
    return out


def navier(case):
    nx, ny, ra, pr, dt, aspect, bc, periodic, init, steps = case
    out = {"params": np.array([nx, ny, ra, pr, dt, aspect, float(periodic), steps]), "bc": np.array(bc), "init": np.array(init)}
    eig = None if periodic else poisson_eig_host(2, nx, 1.0 / aspect ** 2)
    if eig is not None:
        out["poisson_lam"], out["poisson_fwd"], out["poisson_bwd"] = eig
    nav = o.Navier2D(nx, ny, ra, pr, dt, aspect, bc, periodic=periodic, pois_eig=eig)
    if init == "modes":
        nav.set_velocity(0.2, 1.0, 1.0)
        nav.set_temperature(0.2, 1.0, 1.0)
    else:
        nav.init_random(0.1)
    for k, v in nav.state().items():
        out[f"in_{k}"] = v
    for _ in range(steps):
        nav.update()
    for k, v in nav.state().items():
        out[f"out_{k}"] = v
    out["div_norm"] = np.array(nav.div_norm())
    return out


def main():
    for sp in OPERATOR_SPACES:
        np.savez_compressed(os.path.join(HERE, f"operators_{space_name(sp)}.npz"), **operators(sp))
        print("wrote operators", space_name(sp))
    for name, case in NAVIER_CASES.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **navier(case))
        print("wrote", name)


if __name__ == "__main__":
    main()
