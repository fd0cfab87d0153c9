"""Pin the CPU oracle against every golden vector / analytic test the reference
holds for the hot path.  Tolerances: goldens are printed to
4-8 digits in the reference, whose own tolerance is 1e-3 absolute; we check to
the printed precision."""
import numpy as np
import pytest

from oracle import rustpde_oracle as o


def test_dct_convention():
    # [1,2,3,4] -> [2.5, 1.3333, 0, 0.16667]
    c = o.chebyshev(4).forward(np.array([1.0, 2.0, 3.0, 4.0]))
    np.testing.assert_allclose(c, [2.5, 4.0 / 3.0, 0.0, 1.0 / 6.0], atol=1e-14)
    v = o.chebyshev(4).backward(c)
    np.testing.assert_allclose(v, [1, 2, 3, 4], atol=1e-14)


def test_chebyshev_coeffs_are_true_series():
    n = 17
    b = o.chebyshev(n)
    x = b.coords()
    v = 3 * x ** 3 - x + 0.5  # = 0.5 T0 + (9/4 - 1) T1 + 3/4 T3
    c = b.forward(v)
    ref = np.zeros(n)
    ref[0], ref[1], ref[3] = 0.5, 1.25, 0.75
    np.testing.assert_allclose(c, ref, atol=1e-14)
    d = b.differentiate(c, 1)  # 9x^2-1 = 3.5 T0 + 4.5 T2
    ref = np.zeros(n)
    ref[0], ref[2] = 3.5, 4.5
    np.testing.assert_allclose(d, ref, atol=1e-13)


def test_hholtz_adi_1d_golden():
    # src/solver/hholtz_adi.rs:193-212
    nx = 7
    f = o.Field1(o.Space1(o.cheb_dirichlet(nx)))
    h = o.HholtzAdi(f, [1.0])
    x = h.solve(np.arange(1.0, 8.0))
    y = [-0.08214845, -0.10466761, -0.06042153, 0.04809052, 0.04082296]
    np.testing.assert_allclose(x, y, atol=5e-9)


def test_hholtz_adi_2d_golden():
    # src/solver/hholtz_adi.rs:215-246
    nx = 7
    f = o.Field2(o.Space2(o.cheb_dirichlet(nx), o.cheb_dirichlet(nx)))
    h = o.HholtzAdi(f, [1.0, 1.0])
    b = np.tile(np.arange(1.0, 8.0), (nx, 1))
    x = h.solve(b)
    y = np.array([
        [-7.083e-03, -9.025e-03, -5.210e-03, 4.146e-03, 3.520e-03],
        [5.809e-04, 7.402e-04, 4.273e-04, -3.401e-04, -2.887e-04],
        [1.699e-04, 2.165e-04, 1.250e-04, -9.951e-05, -8.447e-05],
        [-1.007e-03, -1.283e-03, -7.406e-04, 5.895e-04, 5.004e-04],
        [-6.775e-04, -8.632e-04, -4.983e-04, 3.966e-04, 3.366e-04],
    ])
    np.testing.assert_allclose(x, y, rtol=6e-4, atol=1e-7)


def test_poisson_1d_golden():
    # src/solver/poisson.rs:275-292
    nx = 8
    f = o.Field1(o.Space1(o.cheb_dirichlet(nx)))
    p = o.Poisson(f, [1.0])
    x = p.solve(np.arange(1.0, 9.0))
    y = [0.1042, 0.0809, 0.0625, 0.0393, -0.0417, -0.0357]
    np.testing.assert_allclose(x, y, atol=6e-5)


GOLD_P2D = np.array([
    [0.01869736, 0.0244178, 0.01403203, -0.0202917, -0.0196697],
    [-0.0027890, -0.004035, -0.0059870, -0.0023490, -0.0046850],
    [-0.0023900, -0.007947, -0.0085570, -0.0189310, -0.0223680],
    [-0.0038940, -0.006622, -0.0096270, -0.0079020, -0.0120490],
    [0.00025400, -0.006752, -0.0082940, -0.0316230, -0.0361640],
    [-0.0001120, -0.004374, -0.0066430, -0.0216410, -0.0262570],
])


def _poisson2d():
    nx, ny = 8, 7
    f = o.Field2(o.Space2(o.cheb_dirichlet(nx), o.cheb_dirichlet(ny)))
    return o.Poisson(f, [1.0, 1.0]), np.tile(np.arange(1.0, 8.0), (nx, 1))


def test_poisson_2d_golden():
    # src/solver/poisson.rs:295-325
    p, b = _poisson2d()
    np.testing.assert_allclose(p.solve(b), GOLD_P2D, atol=1.5e-6)  # rows 1-5 are printed to 6 decimals


def test_poisson_2d_complex_golden():
    # src/solver/poisson.rs:328-361
    p, b = _poisson2d()
    x = p.solve(b * (1 + 1j))
    np.testing.assert_allclose(x, GOLD_P2D * (1 + 1j), atol=1.5e-6)


def test_fdma_tensor_test_matrix():
    # src/solver/fdma_tensor.rs:386-401: the hand-written test matrices are
    # a = laplace_inv_eye . S and c = laplace_inv_eye . B2 . S for cheb_dirichlet(8)
    b = o.cheb_dirichlet(8)
    a = b.laplace_inv_eye() @ b.mass()
    c = b.laplace_inv_eye() @ b.laplace_inv() @ b.mass()
    a_ref = -np.eye(6) + np.eye(6, k=2)
    c_ref = np.array([
        [0.41666, 0.0, -0.2083, 0.0, 0.041666, 0.0],
        [0.0, 0.104166, 0.0, -0.0833, 0.0, 0.0208],
        [-0.0208, 0.0, 0.0542, 0.0, -0.0333, 0.0],
        [0.0, -0.0125, 0.0, 0.033333, 0.0, -0.020833],
        [0.0, 0.0, -0.00833, 0.0, 0.00833, 0.0],
        [0.0, 0.0, 0.0, -0.00595, 0.0, 0.00595],
    ])
    np.testing.assert_allclose(a, a_ref, atol=1e-15)
    np.testing.assert_allclose(c, c_ref, atol=6e-5)
    # and the FdmaTensor residual test of fdma_tensor.rs:376-411
    data = np.arange(36, dtype=float).reshape(6, 6)
    x = o.FdmaTensor([a_ref, a_ref], [c_ref, c_ref], [False, False], 0.0).solve(data)
    np.testing.assert_allclose(a_ref @ x @ c_ref.T + c_ref @ x @ a_ref.T, data, atol=1e-3)


@pytest.mark.parametrize("seed", [0])
def test_fdma_residual(seed):
    # src/solver/fdma.rs:278-305 (residual test M x = b)
    nx = 6
    m = np.zeros((nx, nx))
    for i in range(nx):
        j = i + 1.0
        m[i, i] = 0.5 * j
        if i > 1:
            m[i, i - 2] = 10.0 * j
        if i < nx - 2:
            m[i, i + 2] = 1.5 * j
        if i < nx - 4:
            m[i, i + 4] = 2.5 * j
    data = np.arange(nx, dtype=float)
    x = o.Fdma.from_matrix(m).solve(data, 0)
    np.testing.assert_allclose(m @ x, data, atol=1e-12)
    mv = o.MatVecFdma(np.hstack([m, np.zeros((nx, 2))]))  # matvec.rs:373-404
    d2 = np.arange(nx + 2, dtype=float)
    np.testing.assert_allclose(mv.solve(d2, 0), np.hstack([m, np.zeros((nx, 2))]) @ d2, atol=1e-12)


def test_hholtz_cd_cd_analytic():
    # src/solver/hholtz_adi.rs:249-277
    nx, ny = 16, 7
    f = o.Field2(o.Space2(o.cheb_dirichlet(nx), o.cheb_dirichlet(ny)))
    alpha = 1e-5
    h = o.HholtzAdi(f, [alpha, alpha])
    x, y = f.x
    n = np.pi / 2
    f.v = np.outer(np.cos(n * x), np.cos(n * y))
    expected = f.v / (1 + alpha * n * n * 2)
    f.forward()
    f.vhat = h.solve(f.to_ortho())
    f.backward()
    np.testing.assert_allclose(f.v, expected, atol=1e-3)
    assert np.abs(f.v - expected).max() < 1e-6


def test_hholtz_fo_cd_analytic():
    # src/solver/hholtz_adi.rs:280-308
    nx, ny = 16, 7
    f = o.Field2(o.Space2(o.fourier_r2c(nx), o.cheb_dirichlet(ny)))
    alpha = 1e-5
    h = o.HholtzAdi(f, [alpha, alpha])
    x, y = f.x
    n = np.pi / 2
    f.v = np.outer(np.cos(x), np.cos(n * y))
    expected = f.v / (1 + alpha * n * n + alpha)
    f.forward()
    f.vhat = h.solve(f.to_ortho())
    f.backward()
    assert np.abs(f.v - expected).max() < 1e-6


def test_poisson_cd_cd_analytic():
    # src/solver/poisson.rs:364-393
    nx, ny = 8, 7
    f = o.Field2(o.Space2(o.cheb_dirichlet(nx), o.cheb_dirichlet(ny)))
    p = o.Poisson(f, [1.0, 1.0])
    x, y = f.x
    n = np.pi / 2
    f.v = np.outer(np.cos(n * x), np.cos(n * y))
    expected = -f.v / (n * n * 2)
    f.forward()
    f.vhat = p.solve(f.to_ortho())
    f.backward()
    np.testing.assert_allclose(f.v, expected, atol=1e-3)


def test_poisson_fo_cd_analytic():
    # src/solver/poisson.rs:396-426
    nx, ny = 16, 7
    f = o.Field2(o.Space2(o.fourier_r2c(nx), o.cheb_dirichlet(ny)))
    p = o.Poisson(f, [1.0, 1.0])
    x, y = f.x
    kx, ky = 2.0, np.pi / 2
    f.v = np.outer(np.cos(kx * x), np.cos(ky * y))
    expected = -f.v / (kx * kx + ky * ky)
    f.forward()
    f.vhat = p.solve(f.to_ortho())
    f.backward()
    np.testing.assert_allclose(f.v, expected, atol=1e-3)


@pytest.mark.parametrize("kind", [o.CHEB_DIRICHLET, o.CHEB_NEUMANN])
def test_composite_roundtrip_and_bc(kind):
    n = 33
    b = o.Base(kind, n)
    rng = np.random.default_rng(0)
    c = rng.standard_normal(b.m)
    np.testing.assert_allclose(b.from_ortho(b.to_ortho(c)), c, atol=1e-12)
    v = b.backward(c)
    np.testing.assert_allclose(b.forward(v), c, atol=1e-12)
    if kind == o.CHEB_DIRICHLET:
        assert abs(v[0]) < 1e-12 and abs(v[-1]) < 1e-12
    else:  # Neumann: derivative vanishes at both walls
        dv = o.chebyshev(n).backward(b.differentiate(b.to_ortho(c), 1))
        assert abs(dv[0]) < 1e-9 and abs(dv[-1]) < 1e-9


def test_navier_confined_runs_and_is_divergence_controlled():
    nav = o.Navier2D(33, 33, 1e5, 1.0, 0.01, 1.0, "rbc")
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    for _ in range(20):
        nav.update()
    assert np.isfinite(nav.div_norm()) and nav.div_norm() < 1e-1
    assert all(np.isfinite(v).all() for v in nav.state().values())


def test_navier_periodic_runs():
    nav = o.Navier2D(32, 33, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=True)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    for _ in range(10):
        nav.update()
    assert np.isfinite(nav.div_norm())
    assert nav.state()["temp"].dtype == np.complex128


def test_split_bounds():
    b = o.split_bounds(10, 4)
    assert b == [(0, 1), (2, 3), (4, 6), (7, 9)]
    assert o.split_bounds(8, 2) == [(0, 3), (4, 7)]


def _pdma_test_matrix(n):
    """The matrix of the reference's own solver tests (src/solver/pdma_plus2.rs:209-245, test_pdma_dim1/dim2)."""
    m = np.zeros((n, n))
    for i in range(n):
        j = i + 1.0
        m[i, i] = 0.5 * j
        if i > 1:
            m[i, i - 2] = 10.0 * j
        if i > 0:
            m[i, i - 1] = 4.0 * j
        if i < n - 1:
            m[i, i + 1] = 1.5 * j
        if i < n - 2:
            m[i, i + 2] = 3.5 * j
        if i < n - 3:
            m[i, i + 3] = 4.5 * j
        if i < n - 4:
            m[i, i + 4] = 2.5 * j
    return m


@pytest.mark.parametrize("n", [6, 9, 33])
def test_pdma_plus2_recovers_rhs_like_the_reference_test(n):
    """Row M of SURVEY 8a (bc = "hc", not on the GPU path yet): the oracle's restatement passes the reference's
    test_pdma_dim1 (matrix . solve(data) == data) and agrees with a dense solve."""
    a = _pdma_test_matrix(n)
    data = np.arange(n, dtype=float)
    x = o.PdmaPlus2.from_matrix(a).solve_lane(data)
    np.testing.assert_allclose(a @ x, data, atol=1e-10)
    np.testing.assert_allclose(x, np.linalg.solve(a, data), rtol=1e-10, atol=1e-12)


def test_pdma_plus2_dim2_along_axis0():
    """test_pdma_dim2 (src/solver/pdma_plus2.rs:248-290): lanes along axis 0 of a 6 x 4 array."""
    a = _pdma_test_matrix(6)
    data = np.tile(np.arange(6.0), (4, 1)).T
    x = o.PdmaPlus2(a).solve(data, 0)
    np.testing.assert_allclose(a @ x, data, atol=1e-10)


def test_hholtz_tensor_matches_dense_solve_and_adi_limit():
    """Hholtz (src/solver/hholtz.rs:66-101) restated: the eigendecomposition form must solve the same system as a dense
    direct solve of (C0 x C1 - c0 B0 x C1 - c1 C0 x B1) g = (P0 x P1) f, and reproduce the analytic test of
    hholtz.rs:211-240 ((I - c D2) u = f with u = cos(pi/2 x) cos(pi/2 y) on cd x cd)."""
    from oracle import rustpde_oracle as o

    nx, ny = 18, 14
    fld = o.Field2(o.Space2(o.cheb_dirichlet(nx), o.cheb_dirichlet(ny)))
    c = [0.7, 1.9]
    h = o.Hholtz(fld, c)
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal((nx, ny))
    x = h.solve(rhs)
    a0, b0, p0 = fld.ingredients_for_hholtz(0)
    a1, b1, p1 = fld.ingredients_for_hholtz(1)
    big = np.kron(a0, a1) - c[0] * np.kron(b0, a1) - c[1] * np.kron(a0, b1)
    ref = np.linalg.solve(big, (p0 @ rhs @ p1.T).ravel()).reshape(nx - 2, ny - 2)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-9
    # analytic chain (hholtz.rs:211-240): forward -> to_ortho -> solve -> backward
    n = 64
    fld = o.Field2(o.Space2(o.cheb_dirichlet(n), o.cheb_dirichlet(n)))
    xx, yy = np.meshgrid(fld.x[0], fld.x[1], indexing="ij")
    alpha = 1e-5
    u = np.cos(np.pi / 2 * xx) * np.cos(np.pi / 2 * yy)
    fld.v = (1.0 + alpha * np.pi ** 2 / 2.0) * u
    fld.forward()
    sol = o.Hholtz(fld, [alpha, alpha]).solve(fld.to_ortho())
    fld.vhat = sol
    fld.backward()
    assert np.abs(fld.v - u).max() < 1e-3


def test_fourier_c2c_analytic():
    """FourierC2c (bases.rs:15; funspace semantics unpinned by reference tests): e^{i k x} is mode k with weight n, negative
    wavenumbers sit in the upper half (FFT order), d/dx multiplies by i k, and the Sdma of HholtzAdi divides by 1 + c k^2."""
    n, k = 16, -3
    b = o.fourier_c2c(n)
    x = b.coords()
    v = np.exp(1j * k * x)
    c = b.forward(v)
    want = np.zeros(n, dtype=complex); want[n + k] = n
    np.testing.assert_allclose(c, want, atol=1e-12)
    np.testing.assert_allclose(b.backward(c), v, atol=1e-13)
    np.testing.assert_allclose(b.backward(b.differentiate(c, 1)), 1j * k * v, atol=1e-12)
    assert b.wavenumbers()[n + k] == k and b.laplace()[n + k, n + k] == -k * k
    f = o.Field2(o.Space2(b, o.cheb_dirichlet(9)))
    assert f.v.dtype == np.complex128 and f.vhat.shape == (n, 7)


def _banded_test_matrix(nx, cols=None, cplx=False):
    """the matrix the reference's solver unit tests build (fdma.rs:283-299, :312-332; matvec.rs:376-392)"""
    m = np.zeros((nx, cols or nx), dtype=np.complex128 if cplx else np.float64)
    for i in range(nx):
        j = i + 1.0
        m[i, i] = 0.5 * j + (1.5j * j if cplx else 0)
        if i > 1:
            m[i, i - 2] = 10.0 * j + (12.0j * j if cplx else 0)
        if i < nx - 2:
            m[i, i + 2] = 1.5 * j + (4.5j * j if cplx else 0)
        if i < nx - 4:
            m[i, i + 4] = 2.5 * j
    return m


def test_matvecfdma_dim2_like_the_reference_test():
    # src/solver/matvec.rs:372-404: (nx, nx + 2) banded matrix against a dense dot along either axis of a 2-D array
    nx = 6
    m = _banded_test_matrix(nx, nx + 2)
    data = np.arange((nx + 2) ** 2, dtype=float).reshape(nx + 2, nx + 2)
    mv = o.MatVecFdma(m)
    np.testing.assert_allclose(mv.solve(data, 0), m @ data, atol=1e-3)        # approx_eq of the reference: 1e-3 absolute
    np.testing.assert_allclose(mv.solve(data, 1), (m @ data.T).T, atol=1e-3)
    np.testing.assert_allclose(mv.solve(data, 0), m @ data, rtol=1e-14)
    np.testing.assert_allclose(mv.solve(data, 1), (m @ data.T).T, rtol=1e-14)


def test_fdma_dim1_complex_like_the_reference_test():
    # src/solver/fdma.rs:306-337: complex matrix and right-hand side, M x recovers the data
    nx = 6
    m = _banded_test_matrix(nx, cplx=True)
    data = np.arange(nx) + 1j * (np.arange(nx) + 1.0)
    low, dia, up1, up2 = (np.diagonal(m, k).copy() for k in (-2, 0, 2, 4))
    # the reference's Fdma is generic over the scalar; the oracle's restatement is real (every Navier2D system is real), so the
    # complex case is checked through the same sweep / elimination written out on complex diagonals
    n = nx
    for i in range(2, n):          # fdma.rs:73-82
        low[i - 2] /= dia[i - 2]
        dia[i] -= low[i - 2] * up1[i - 2]
        if i < n - 2:
            up1[i] -= low[i - 2] * up2[i - 2]
    f = o.Fdma(np.zeros(n - 2), np.ones(n), np.zeros(n - 2), np.zeros(n - 4), sweep=False)
    f.low, f.dia, f.up1, f.up2, f.sweeped = low, dia, up1, up2, True
    x = data.copy()
    f.fdma(x)                      # fdma.rs:101-118 on complex values
    np.testing.assert_allclose(m @ x, data, atol=1e-12)


def test_sdma_like_the_reference_tests():
    # src/solver/sdma.rs:137-175: diagonal matrix, real and complex
    nx = 6
    m = np.diag(0.5 * (np.arange(nx) + 1.0))
    data = np.arange(nx, dtype=float)
    np.testing.assert_allclose(m @ o.Sdma(m).solve(data, 0), data, atol=1e-14)
    mc = np.diag((0.5 + 0.5j) * (np.arange(nx) + 1.0))
    dc = np.arange(nx) + 1j * (np.arange(nx) + 1.0)
    np.testing.assert_allclose(mc @ o.Sdma(mc).solve(dc, 0), dc, atol=1e-14)
    d2 = np.arange(nx * 4, dtype=float).reshape(4, nx)   # along axis 1 of a 2-D array
    np.testing.assert_allclose(o.Sdma(m).solve(d2, 1) @ m.T, d2, atol=1e-13)


def test_average_doctest_goldens():
    # src/field/average.rs:12-25 and :38-52 (doc tests): Chebyshev 6 x 5, v[i, j] = j  =>  average_axis(0) = [0, 1, 2, 3, 4],
    # average() = 2 -- pins the dx weights of Field2 (src/field.rs:135-163) and both reductions
    f = o.Field2(o.Space2(o.chebyshev(6), o.chebyshev(5)))
    f.v = np.tile(np.arange(5.0), (6, 1))
    np.testing.assert_allclose(o.Navier2D.average_axis(f, 0), np.arange(5.0), rtol=0, atol=2e-15)
    assert abs(o.Navier2D.average(f) - 2.0) < 2e-15


def test_eig_like_the_reference_test():
    # src/solver/utils.rs:183-204: Q diag(lam) Q^-1 reproduces the matrix (1e-3 absolute in the reference); eigenvalues descending
    a = np.tile(np.arange(1.0, 6.0), (5, 1))
    lam, q, qinv = o.eig(a)
    np.testing.assert_allclose(q @ np.diag(lam) @ qinv, a, atol=1e-3)
    np.testing.assert_allclose(q @ np.diag(lam) @ qinv, a, atol=1e-12)
    assert np.all(np.diff(lam) <= 1e-12) and abs(lam[0] - 15.0) < 1e-12
