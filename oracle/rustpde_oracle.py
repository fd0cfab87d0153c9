"""CPU oracle for the Navier2D / Navier2DMpi per-timestep spectral hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product path (``rustpde_mpi_b200``) never does.

This is a numpy/scipy f64 *restatement* of the reference's algorithm (the
reference is Rust and cannot be built in this image: no cargo/rustc, and the
transform arithmetic lives in the un-vendored crate ``funspace 0.3.0``,
``/root/reference/Cargo.toml:17``, ``Cargo.lock:451-465``).  Every function
cites the reference file:line it follows (paths relative to /root/reference).

Parity pin: ``tests/test_oracle_golden.py`` checks this oracle against every
golden vector the reference's own tests hold for the path
(``src/solver/hholtz_adi.rs:193-246``, ``src/solver/poisson.rs:275-361``,
``src/solver/fdma_tensor.rs:394-401``) and its analytic end-to-end tests
(``hholtz_adi.rs:249-308``, ``poisson.rs:364-426``).  The funspace-side
conventions that no reference test pins (forward normalisation, r2c scaling,
ChebNeumann stencil, differentiation) follow SURVEY.md Appendix A and are
flagged "unpinned" there and in DESIGN.md.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
from scipy.fft import dct, fft, ifft, irfft, rfft

# BaseKind enum order, src/field.rs:173-177 (funspace ``BaseKind``)
CHEBYSHEV, CHEB_DIRICHLET, CHEB_NEUMANN, CHEB_DIRICHLET_NEUMANN, FOURIER_R2C, FOURIER_C2C = range(6)
_COMPOSITE = (CHEB_DIRICHLET, CHEB_NEUMANN, CHEB_DIRICHLET_NEUMANN)
_CHEB = (CHEBYSHEV,) + _COMPOSITE


def _mv(a, axis):
    """Bring ``axis`` to the front (ops below act on axis 0)."""
    return np.moveaxis(a, axis, 0)


# ---------------------------------------------------------------------------
# L1: bases (funspace 0.3.0; call sites src/bases.rs:11-19, src/field.rs:104-128)
# ---------------------------------------------------------------------------
class Base:
    """One 1-D basis: chebyshev / cheb_dirichlet / cheb_neumann /
    cheb_dirichlet_neumann / fourier_r2c (src/bases.rs:11-19)."""

    def __init__(self, kind: int, n: int):
        self.kind, self.n = kind, n
        if kind == CHEBYSHEV:
            self.m = n
        elif kind in _COMPOSITE:
            self.m = n - 2
        elif kind == FOURIER_R2C:
            self.m = n // 2 + 1
        elif kind == FOURIER_C2C:
            self.m = n   # bases.rs:15: complex in, complex out, n modes in FFT order (not on the Navier2D path; funspace semantics unpinned)
        else:
            raise ValueError("bad base kind")
        self.is_cheb = kind in _CHEB

    # -- grid -------------------------------------------------------------
    def coords(self):
        """Gauss-Lobatto nodes, ascending / equispaced (A.4)."""
        n = self.n
        if self.is_cheb:
            return -np.cos(np.pi * np.arange(n) / (n - 1))
        return 2.0 * np.pi * np.arange(n) / n

    def wavenumbers(self):
        if self.kind == FOURIER_C2C:
            return np.fft.fftfreq(self.n, 1.0 / self.n)   # 0 .. n/2-1, -n/2 .. -1
        return np.arange(self.m, dtype=np.float64)

    # -- stencil -----------------------------------------------
    def stencil_coeffs(self):
        """(s1, s2): ortho_k = c_k + s1[k-1] c_{k-1} + s2[k-2] c_{k-2}."""
        m = self.m
        k = np.arange(m, dtype=np.float64)
        if self.kind == CHEB_DIRICHLET:
            return np.zeros(m), -np.ones(m)
        if self.kind == CHEB_NEUMANN:
            return np.zeros(m), -((k / (k + 2.0)) ** 2)
        if self.kind == CHEB_DIRICHLET_NEUMANN:
            a = ((k + 2.0) ** 2 - k ** 2) / ((k + 1.0) ** 2 + (k + 2.0) ** 2)
            return a, a - 1.0
        return None

    def stencil(self):
        """Dense S (n x m); identity for orthogonal bases.  funspace ``mass``."""
        if self.kind not in _COMPOSITE:
            return np.eye(self.m)
        s1, s2 = self.stencil_coeffs()
        S = np.zeros((self.n, self.m))
        idx = np.arange(self.m)
        S[idx, idx] = 1.0
        S[idx + 1, idx] = s1
        S[idx + 2, idx] = s2
        return S

    # -- transforms along axis 0 ---------------------------------------------
    def _cheb_fwd(self, v):
        n = self.n
        sign = np.where(np.arange(n) % 2 == 0, 1.0, -1.0).reshape((n,) + (1,) * (v.ndim - 1))
        c = dct(v, type=1, axis=0) * sign / (n - 1)
        c[0] *= 0.5
        c[n - 1] *= 0.5
        return c

    def _cheb_bwd(self, c):
        n = self.n
        sign = np.where(np.arange(n) % 2 == 0, 1.0, -1.0).reshape((n,) + (1,) * (c.ndim - 1))
        y = c * sign
        y[0] *= 2.0
        y[n - 1] *= 2.0
        return dct(y, type=1, axis=0) * 0.5

    def to_ortho(self, c, axis=0):
        """composite -> orthonormal, src/field.rs:113-115 (funspace to_ortho)."""
        if self.kind not in _COMPOSITE:
            return np.array(c, copy=True)
        c = _mv(c, axis)
        s1, s2 = self.stencil_coeffs()
        sh = (self.m,) + (1,) * (c.ndim - 1)
        o = np.zeros((self.n,) + c.shape[1:], dtype=c.dtype)
        o[: self.m] += c
        if self.kind == CHEB_DIRICHLET_NEUMANN:
            o[1 : self.m + 1] += s1.reshape(sh) * c
        o[2:] += s2.reshape(sh) * c
        return np.moveaxis(o, 0, axis)

    def from_ortho(self, o, axis=0):
        """orthonormal -> composite: c = (S^T S)^-1 S^T o, src/field.rs:118-123."""
        if self.kind not in _COMPOSITE:
            return np.array(o, copy=True)
        o = _mv(o, axis)
        S = self.stencil()
        rhs = np.tensordot(S.T, o, axes=(1, 0))
        sts = S.T @ S
        m = self.m
        ab = np.zeros((5, m))
        for off in range(-2, 3):
            d = np.diagonal(sts, off)
            if off >= 0:
                ab[2 - off, off:] = d
            else:
                ab[2 - off, : m + off] = d
        flat = rhs.reshape(m, -1)
        if np.iscomplexobj(flat):
            sol = sla.solve_banded((2, 2), ab, flat.real) + 1j * sla.solve_banded((2, 2), ab, flat.imag)
        else:
            sol = sla.solve_banded((2, 2), ab, flat)
        return np.moveaxis(sol.reshape(rhs.shape), 0, axis)

    def forward(self, v, axis=0):
        """physical -> spectral along axis, src/field.rs:103-105."""
        v = _mv(v, axis)
        if self.is_cheb:
            c = self._cheb_fwd(np.asarray(v))
            if self.kind in _COMPOSITE:
                c = self.from_ortho(c, 0)
        elif self.kind == FOURIER_C2C:
            c = fft(np.asarray(v, dtype=np.complex128), axis=0)
        else:
            c = rfft(v, axis=0)
        return np.moveaxis(c, 0, axis)

    def backward(self, c, axis=0):
        """spectral -> physical along axis, src/field.rs:108-110."""
        c = _mv(c, axis)
        if self.is_cheb:
            v = self._cheb_bwd(self.to_ortho(c, 0))
        elif self.kind == FOURIER_C2C:
            v = ifft(c, axis=0)
        else:
            v = irfft(c, n=self.n, axis=0)
        return np.moveaxis(v, 0, axis)

    def differentiate(self, o, d, axis=0):
        """d derivatives of *orthonormal* coefficients."""
        o = np.array(_mv(o, axis), copy=True)
        if d == 0:
            return np.moveaxis(o, 0, axis)
        if self.is_cheb:
            n = self.n
            for _ in range(d):
                b = np.zeros_like(o)
                # b_k = b_{k+2} + 2(k+1) a_{k+1}, k descending; b_0 /= 2
                for k in range(n - 2, -1, -1):
                    b[k] = 2.0 * (k + 1) * o[k + 1]
                    if k + 2 < n:
                        b[k] += b[k + 2]
                b[0] *= 0.5
                o = b
        else:
            k = self.wavenumbers().reshape((self.m,) + (1,) * (o.ndim - 1))
            o = o * (1j * k) ** d
        return np.moveaxis(o, 0, axis)

    # -- matrices (funspace mass / laplace / laplace_inv / laplace_inv_eye) -----
    def mass(self):
        return self.stencil()

    def laplace(self):
        if self.kind in (FOURIER_R2C, FOURIER_C2C):
            return np.diag(-self.wavenumbers() ** 2)
        raise NotImplementedError("laplace() is only read for Fourier axes (src/field.rs:213)")

    def laplace_inv(self):
        """B2: quasi-inverse of D^2 (n x n), rows 0,1 zero."""
        n = self.n
        B = np.zeros((n, n))
        for i in range(2, n):
            B[i, i - 2] = 0.25 if i == 2 else 1.0 / (4.0 * i * (i - 1))
            if i < n - 2:
                B[i, i] = -1.0 / (2.0 * (i * i - 1.0))
            if i < n - 4:
                B[i, i + 2] = 1.0 / (4.0 * i * (i + 1))
        return B

    def laplace_inv_eye(self):
        return np.eye(self.n)[2:, :]


def chebyshev(n):
    return Base(CHEBYSHEV, n)


def cheb_dirichlet(n):
    return Base(CHEB_DIRICHLET, n)


def cheb_neumann(n):
    return Base(CHEB_NEUMANN, n)


def cheb_dirichlet_neumann(n):
    return Base(CHEB_DIRICHLET_NEUMANN, n)


def fourier_r2c(n):
    return Base(FOURIER_R2C, n)


def fourier_c2c(n):
    return Base(FOURIER_C2C, n)


class Space2:
    """funspace ``Space2`` as used through src/field.rs:81-129."""

    def __init__(self, b0: Base, b1: Base):
        self.bases = (b0, b1)

    def base_kind(self, axis):
        return self.bases[axis].kind

    def shape_physical(self):
        return (self.bases[0].n, self.bases[1].n)

    def shape_spectral(self):
        return (self.bases[0].m, self.bases[1].m)

    def spectral_dtype(self):
        return np.complex128 if self.bases[0].kind in (FOURIER_R2C, FOURIER_C2C) else np.float64

    def ndarray_physical(self):
        return np.zeros(self.shape_physical(), dtype=np.complex128 if self.bases[0].kind == FOURIER_C2C else np.float64)

    def ndarray_spectral(self):
        return np.zeros(self.shape_spectral(), dtype=self.spectral_dtype())

    def coords(self):
        return [b.coords() for b in self.bases]

    def forward(self, v):
        # r2c x cheb: real axis 1 first, then axis 0
        return self.bases[0].forward(self.bases[1].forward(v, 1), 0)

    def backward(self, vhat):
        return self.bases[1].backward(self.bases[0].backward(vhat, 0), 1)

    def to_ortho(self, vhat):
        return self.bases[1].to_ortho(self.bases[0].to_ortho(vhat, 0), 1)

    def from_ortho(self, o):
        return self.bases[1].from_ortho(self.bases[0].from_ortho(o, 0), 1)

    def gradient(self, vhat, deriv, scale=None):
        """src/field.rs:127-129: to_ortho, differentiate per axis, / scale^d."""
        o = self.to_ortho(vhat)
        for ax in (0, 1):
            o = self.bases[ax].differentiate(o, deriv[ax], ax)
        if scale is not None:
            o = o / (scale[0] ** deriv[0] * scale[1] ** deriv[1])
        return o

    def mass(self, axis):
        return self.bases[axis].mass()

    def laplace(self, axis):
        return self.bases[axis].laplace()

    def laplace_inv(self, axis):
        return self.bases[axis].laplace_inv()

    def laplace_inv_eye(self, axis):
        return self.bases[axis].laplace_inv_eye()


# ---------------------------------------------------------------------------
# L2: Field (src/field.rs)
# ---------------------------------------------------------------------------
class Field2:
    """``FieldBase`` for N = 2, src/field.rs:59-129."""

    def __init__(self, space: Space2):
        self.space = space
        self.v = space.ndarray_physical()
        self.vhat = space.ndarray_spectral()
        self.x = space.coords()
        self.dx = [self._get_dx(x, space.bases[i].kind in (FOURIER_R2C, FOURIER_C2C)) for i, x in enumerate(self.x)]

    @staticmethod
    def _get_dx(x, periodic):
        """src/field.rs:135-163."""
        if periodic:
            return np.full(len(x), x[2] - x[1])
        left = np.concatenate(([x[0]], 0.5 * (x[1:] + x[:-1])))
        right = np.concatenate((0.5 * (x[1:] + x[:-1]), [x[-1]]))
        return right - left

    def scale(self, scale):
        """src/field.rs:93-100."""
        for i, sc in enumerate(scale):
            self.x[i] = self.x[i] * sc
            self.dx[i] = self.dx[i] * sc

    def forward(self):
        self.vhat = self.space.forward(self.v)

    def backward(self):
        self.v = self.space.backward(self.vhat)

    def to_ortho(self):
        return self.space.to_ortho(self.vhat)

    def from_ortho(self, o):
        self.vhat = self.space.from_ortho(o)

    def gradient(self, deriv, scale=None):
        return self.space.gradient(self.vhat, deriv, scale)

    def ingredients_for_hholtz(self, axis):
        """src/field.rs:195-216."""
        b = self.space.bases[axis]
        mass = b.mass()
        if b.kind == CHEBYSHEV:
            peye = b.laplace_inv_eye()
            pinv = peye @ b.laplace_inv()
            ms = mass[:, 2:]
            return pinv @ ms, peye @ ms, pinv
        if b.kind in _COMPOSITE:
            peye = b.laplace_inv_eye()
            pinv = peye @ b.laplace_inv()
            return pinv @ mass, peye @ mass, pinv
        return mass, b.laplace(), None

    def ingredients_for_poisson(self, axis):
        """src/field.rs:229-249."""
        a, bm, pre = self.ingredients_for_hholtz(axis)
        return a, bm, pre, self.space.bases[axis].kind in (FOURIER_R2C, FOURIER_C2C)   # src/field.rs:244


# ---------------------------------------------------------------------------
# L3: solvers (src/solver/*.rs) -- vectorised over lanes, sequential along the lane
# ---------------------------------------------------------------------------
def diag(a, offset):
    """src/solver/utils.rs:17-44."""
    return np.array(np.diagonal(a, offset), copy=True)


class Sdma:
    """src/solver/sdma.rs:19-46."""

    def __init__(self, a):
        self.n = a.shape[0]
        self.dia = diag(a, 0)

    def solve(self, inp, axis):
        x = _mv(np.array(inp, copy=True), axis)
        x /= self.dia.reshape((self.n,) + (1,) * (x.ndim - 1))
        return np.moveaxis(x, 0, axis)


class Fdma:
    """4-diagonal (-2,0,2,4) solver, src/solver/fdma.rs:33-118."""

    def __init__(self, low, dia, up1, up2, sweep=True):
        self.n = len(dia)
        self.low, self.dia, self.up1, self.up2 = (np.array(v, dtype=np.float64, copy=True) for v in (low, dia, up1, up2))
        self.sweeped = False
        if sweep:
            self.sweep()

    @classmethod
    def from_matrix(cls, a):
        return cls(diag(a, -2), diag(a, 0), diag(a, 2), diag(a, 4), sweep=True)

    @classmethod
    def from_matrix_raw(cls, a):
        return cls(diag(a, -2), diag(a, 0), diag(a, 2), diag(a, 4), sweep=False)

    def sweep(self):
        """src/solver/fdma.rs:73-82."""
        n = self.n
        for i in range(2, n):
            self.low[i - 2] /= self.dia[i - 2]
            self.dia[i] -= self.low[i - 2] * self.up1[i - 2]
            if i < n - 2:
                self.up1[i] -= self.low[i - 2] * self.up2[i - 2]
        self.sweeped = True

    def fdma(self, x):
        """src/solver/fdma.rs:101-118; x has the lane on axis 0, modified in place."""
        n = self.n
        low, dia, up1, up2 = self.low, self.dia, self.up1, self.up2
        for i in range(2, n):
            x[i] = x[i] - x[i - 2] * low[i - 2]
        x[n - 1] = x[n - 1] / dia[n - 1]
        x[n - 2] = x[n - 2] / dia[n - 2]
        x[n - 3] = (x[n - 3] - x[n - 1] * up1[n - 3]) / dia[n - 3]
        x[n - 4] = (x[n - 4] - x[n - 2] * up1[n - 4]) / dia[n - 4]
        for i in range(n - 5, -1, -1):
            x[i] = (x[i] - x[i + 2] * up1[i] - x[i + 4] * up2[i]) / dia[i]

    def solve(self, inp, axis):
        assert self.sweeped, "Fdma: Forward sweep must be performed before solve!"
        x = np.array(_mv(inp, axis), copy=True)
        assert x.shape[0] == self.n
        self.fdma(x)
        return np.moveaxis(x, 0, axis)

    def add_scaled(self, other, lam):
        """&self + &(&other * lam), src/solver/fdma.rs:195-243 (unsweeped only)."""
        assert not self.sweeped and not other.sweeped
        return Fdma(self.low + other.low * lam, self.dia + other.dia * lam,
                    self.up1 + other.up1 * lam, self.up2 + other.up2 * lam, sweep=False)


class PdmaPlus2:
    """Banded solve for diagonals at offsets -2..+4 (``PdmaPlus2``, src/solver/pdma_plus2.rs:19-157; used by the
    ChebDirichletNeumann bases of bc = "hc", SURVEY 8a row M -- not on the GPU path yet).  Restated as what it is:
    LU elimination without pivoting restricted to the band (two sub-diagonals, four super-diagonals), the factor
    rows normalised by the pivot like the reference's al/be/ga/de, followed by forward and backward substitution."""

    LOW, UP = 2, 4

    def __init__(self, a):
        a = np.asarray(a, dtype=float)
        n = a.shape[0]
        self.n = n
        u = a.copy()                       # becomes the upper factor (rows divided by their pivot at the end)
        low = np.zeros((n, self.LOW))      # multipliers of the rows below the pivot
        for k in range(n):
            for i in range(k + 1, min(n, k + self.LOW + 1)):
                m = u[i, k] / u[k, k]
                low[i, i - k - 1] = m
                hi = min(n, k + self.UP + 1)
                u[i, k:hi] -= m * u[k, k:hi]
        self.piv = np.diag(u).copy()       # mu
        self.up = np.zeros((n, self.UP))   # al, be, ga, de
        for k in range(n):
            for j in range(1, self.UP + 1):
                if k + j < n:
                    self.up[k, j - 1] = u[k, k + j] / self.piv[k]
        self.low = low

    @classmethod
    def from_matrix(cls, a):
        return cls(a)

    def solve_lane(self, rhs):
        n = self.n
        z = np.array(rhs, dtype=np.result_type(rhs, float), copy=True)
        for i in range(n):                 # forward: L z = rhs, then scale by the pivot
            for j in range(1, self.LOW + 1):
                if i - j >= 0:
                    z[i] -= self.low[i, j - 1] * z[i - j] * self.piv[i - j]
            z[i] = z[i] / self.piv[i]
        x = z
        for i in range(n - 1, -1, -1):     # backward with the normalised upper factor
            for j in range(1, self.UP + 1):
                if i + j < n:
                    x[i] -= self.up[i, j - 1] * x[i + j]
        return x

    def solve(self, inp, axis):
        return np.apply_along_axis(self.solve_lane, axis, inp)


class MatVecFdma:
    """banded (n-2) x n mat-vec, src/solver/matvec.rs:161-228."""

    def __init__(self, a):
        m, n = a.shape
        self.m, self.n = m, n
        self.low = np.zeros(m)
        self.dia = np.zeros(m)
        self.up1 = np.zeros(m)
        self.up2 = np.zeros(m)
        for i in range(m):
            self.dia[i] = a[i, i]
            if i > 1:
                self.low[i] = a[i, i - 2]
            if i < m - 2:
                self.up1[i] = a[i, i + 2]
            if i < m - 4:
                self.up2[i] = a[i, i + 4]

    def solve(self, inp, axis):
        x = _mv(inp, axis)
        m = self.m
        sh = (m,) + (1,) * (x.ndim - 1)
        out = x[:m] * self.dia.reshape(sh)
        out[2:] += x[: m - 2] * self.low[2:].reshape((m - 2,) + sh[1:])
        out[: m - 2] += x[2:m] * self.up1[: m - 2].reshape((m - 2,) + sh[1:])
        out[: m - 4] += x[4:m] * self.up2[: m - 4].reshape((m - 4,) + sh[1:])
        return np.moveaxis(out, 0, axis)


def eig(a):
    """src/solver/utils.rs:67-100: LAPACK real eig, real parts, sort descending."""
    ev, evec = sla.eig(a)
    ev, evec = ev.real, evec.real
    perm = np.argsort(ev, kind="stable")[::-1]
    ev, evec = ev[perm], evec[:, perm]
    return ev, evec, np.linalg.inv(evec)


def parity_eig(a, c):
    """Same decomposition as ``FdmaTensor.from_matrix`` (src/solver/fdma_tensor.rs:117-129) of X = C^-1 A, computed
    on the two parity blocks separately: A and C only couple indices of equal parity, so X is block diagonal after
    an even/odd permutation and two half-size LAPACK problems replace the full one (8x less work at n = 4095; the
    solve x = Q (..) Q^-1 C^-1 rhs does not depend on how eigenvectors are scaled or grouped).  Used by bench.py
    for the large configurations only; parity tests hand one decomposition to both sides anyway."""
    m = a.shape[0]
    lam = np.zeros(m)
    q = np.zeros((m, m))
    fwd = np.zeros((m, m))
    for par in (0, 1):
        idx = np.arange(par, m, 2)
        cinv = np.linalg.inv(c[np.ix_(idx, idx)])
        l_p, q_p, p_p = eig(cinv @ a[np.ix_(idx, idx)])
        lam[idx] = l_p
        q[np.ix_(idx, idx)] = q_p
        fwd[np.ix_(idx, idx)] = p_p @ cinv
    perm = np.argsort(lam, kind="stable")[::-1]
    return lam[perm], fwd[perm, :], q[:, perm]


class FdmaTensor:
    """src/solver/fdma_tensor.rs:74-154 (N = 1, 2)."""

    def __init__(self, a, c, a_is_diag, alpha=0.0):
        ndim = len(a)
        self.fwd, self.bwd, self.lam = [], [], []
        for i in range(ndim - 1):
            if a_is_diag[i]:
                self.lam.append(diag(a[i], 0))
                self.fwd.append(None)
                self.bwd.append(None)
            else:
                cinv = np.linalg.inv(c[i])
                lam, q, p = eig(cinv @ a[i])
                self.lam.append(lam)
                self.fwd.append(p @ cinv)
                self.bwd.append(q)
        self.n = a[-1].shape[0]
        self.fdma = [Fdma.from_matrix_raw(a[-1]), Fdma.from_matrix_raw(c[-1])]
        self.alpha = alpha
        self.ndim = ndim
        if ndim == 1:
            self.fdma[0].sweep()

    def solve(self, inp):
        """src/solver/fdma_tensor.rs:236-290."""
        if self.ndim == 1:
            return self.fdma[0].solve(inp, 0)
        assert inp.shape[0] == len(self.lam[0]) and inp.shape[1] == self.n, "Dimension mismatch in Tensor!"
        out = self.fwd[0] @ inp if self.fwd[0] is not None else np.array(inp, copy=True)
        res = np.empty_like(out)
        for i, lam in enumerate(self.lam[0]):
            f = self.fdma[0].add_scaled(self.fdma[1], lam + self.alpha)
            f.sweep()
            res[i] = f.solve(out[i], 0)
        if self.bwd[0] is not None:
            res = self.bwd[0] @ res
        return res


class HholtzAdi:
    """src/solver/hholtz_adi.rs:36-169:  (I - c D2) vhat = f by ADI."""

    def __init__(self, field, c):
        self.solver, self.matvec = [], []
        nd = len(c)
        for axis in range(nd):
            mat_a, mat_b, pre = field.ingredients_for_hholtz(axis) if nd == 2 else _ingredients_1d(field, axis)
            mat = mat_a - mat_b * c[axis]
            kind = field.space.bases[axis].kind
            if kind in (CHEBYSHEV, CHEB_DIRICHLET, CHEB_NEUMANN):
                self.solver.append(Fdma.from_matrix(mat))
            elif kind == CHEB_DIRICHLET_NEUMANN:
                self.solver.append(PdmaPlus2.from_matrix(mat))   # hholtz_adi.rs:64
            else:
                self.solver.append(Sdma(mat))
            self.matvec.append(MatVecFdma(pre) if pre is not None else None)

    def solve(self, inp):
        rhs = inp
        nd = len(self.solver)
        for ax in range(nd):
            if self.matvec[ax] is not None:
                rhs = self.matvec[ax].solve(rhs, ax)
        out = rhs
        for ax in range(nd):
            out = self.solver[ax].solve(out, ax)
        return out


class Poisson:
    """src/solver/poisson.rs:42-236: c D2 vhat = f via eigendecomposition of axis 0."""

    def __init__(self, field, c, eig=None):
        """``eig`` = (lam, fwd, bwd) replaces the LAPACK eigendecomposition of axis 0 (lam already
        carries the singularity shift).  Parity tests pass the SAME decomposition to the oracle and
        to the CUDA solver: the shifted-singular mode amplifies LAPACK-build rounding differences
        to ~1e-9, which would otherwise mask the 1e-10 comparison ("Poisson parity")."""
        nd = len(c)
        lap, mass, isd, self.matvec = [], [], [], []
        for axis in range(nd):
            mat_a, mat_b, pre, is_diag = field.ingredients_for_poisson(axis) if nd == 2 else _ingredients_1d(field, axis) + (False,)
            mass.append(mat_a)
            lap.append(mat_b * c[axis])
            self.matvec.append(MatVecFdma(pre) if pre is not None else None)
            isd.append(is_diag)
        if isinstance(eig, str) and eig == "parity":
            if isd[0]:
                eig = None
            else:
                lam_p, fwd_p, bwd_p = parity_eig(lap[0], mass[0])
                if abs(lam_p[0]) < 1e-10:   # singularity hack, src/solver/poisson.rs:84-86
                    lam_p = lam_p - 1e-10
                eig = (lam_p, fwd_p, bwd_p)
        if eig is not None and not isd[0]:
            t = FdmaTensor.__new__(FdmaTensor)
            t.ndim, t.alpha, t.n = 2, 0.0, lap[-1].shape[0]
            t.lam = [np.array(eig[0], copy=True)]
            t.fwd = [np.array(eig[1], copy=True)]
            t.bwd = [np.array(eig[2], copy=True)]
            t.fdma = [Fdma.from_matrix_raw(lap[-1]), Fdma.from_matrix_raw(mass[-1])]
            self.solver = t
            return
        self.solver = FdmaTensor(lap, mass, isd, 0.0)
        # singularity hack, src/solver/poisson.rs:84-86 (shifts the WHOLE lam[0] array)
        if nd == 2 and abs(self.solver.lam[0][0]) < 1e-10:
            self.solver.lam[0] = self.solver.lam[0] - 1e-10

    def solve(self, inp):
        rhs = inp
        for ax in range(len(self.matvec)):
            if self.matvec[ax] is not None:
                rhs = self.matvec[ax].solve(rhs, ax)
        return self.solver.solve(rhs)


class Hholtz:
    """src/solver/hholtz.rs:66-101,153-176: (I - c D2) vhat = A f via the eigendecomposition of axis 0 (FdmaTensor, alpha = 1)."""

    def __init__(self, field, c, eig=None):
        nd = len(c)
        lap, mass, isd, self.matvec = [], [], [], []
        for axis in range(nd):
            mat_a, mat_b, pre, is_diag = field.ingredients_for_poisson(axis) if nd == 2 else _ingredients_1d(field, axis) + (False,)
            mass.append(mat_a)
            lap.append(-1.0 * mat_b * c[axis])
            self.matvec.append(MatVecFdma(pre) if pre is not None else None)
            isd.append(is_diag)
        if eig is not None and nd == 2 and not isd[0]:
            t = FdmaTensor.__new__(FdmaTensor)
            t.ndim, t.alpha, t.n = 2, 1.0, lap[-1].shape[0]
            t.lam = [np.array(eig[0], copy=True)]
            t.fwd = [np.array(eig[1], copy=True)]
            t.bwd = [np.array(eig[2], copy=True)]
            t.fdma = [Fdma.from_matrix_raw(lap[-1]), Fdma.from_matrix_raw(mass[-1])]
            self.solver = t
            return
        self.solver = FdmaTensor(lap, mass, isd, 1.0)

    def solve(self, inp):
        rhs = inp
        for ax in range(len(self.matvec)):
            if self.matvec[ax] is not None:
                rhs = self.matvec[ax].solve(rhs, ax)
        return self.solver.solve(rhs)


class Space1:
    def __init__(self, b0):
        self.bases = (b0,)


class Field1:
    def __init__(self, space):
        self.space = space


def _ingredients_1d(field, axis):
    f2 = Field2.__new__(Field2)
    f2.space = Space2(field.space.bases[0], field.space.bases[0])
    return Field2.ingredients_for_hholtz(f2, 0)


# ---------------------------------------------------------------------------
# L4: Navier2D (src/navier_stokes/*.rs)
# ---------------------------------------------------------------------------
def get_nu(ra, pr, height):
    """src/navier_stokes/functions.rs:12-15."""
    return np.sqrt(pr / (ra / height ** 3.0))


def get_ka(ra, pr, height):
    """src/navier_stokes/functions.rs:18-21."""
    return np.sqrt(1.0 / ((ra / height ** 3.0) * pr))


def dealias(vhat):
    """2/3 rule, src/navier_stokes/functions.rs:72-82 (bit-exact index rule)."""
    n_x = vhat.shape[0] * 2 // 3
    n_y = vhat.shape[1] * 2 // 3
    vhat[n_x:, :] = 0
    vhat[:, n_y:] = 0


def bc_rbc(b0, ny):
    """src/navier_stokes/boundary_conditions.rs:18-36 / :143-161 (periodic)."""
    f = Field2(Space2(b0, chebyshev(ny)))
    x = f.x[1]
    x1, x2 = x[0], x[-1]
    y1, y2 = 0.5, -0.5
    m = (y2 - y1) / (x2 - x1)
    n = (y1 * x2 - y2 * x1) / (x2 - x1)
    f.v[:, :] = (m * x + n)[None, :]
    f.forward()
    f.backward()
    return f


def bc_hc(b0, ny, periodic=False):
    """src/navier_stokes/boundary_conditions.rs:103-135 (confined) / :165-195 (periodic): T = -0.5 cos(2 pi (x - x0) / L)
    at the bottom, T = T' = 0 at the top, a parabola in y with its vertex at the top wall."""
    f = Field2(Space2(b0, chebyshev(ny)))
    x, y = f.x
    x0, length = x[0], x[-1] - x[0]   # (periodic: the Fourier grid stops one point short of 2 pi; the reference uses x[last] all the same)
    f_x = -0.5 * np.cos(2.0 * np.pi * (x - x0) / length)
    yl, yr = y[0], y[-1]
    f.v[:, :] = (f_x / (yl - yr) ** 2)[:, None] * ((y - yr) ** 2)[None, :]
    f.forward()
    f.backward()
    return f


class Navier2D:
    """``Navier2D`` with bc = "rbc" or "hc", src/navier_stokes/navier.rs:49-466."""

    def __init__(self, nx, ny, ra, pr, dt, aspect, bc="rbc", periodic=False, pois_eig=None):
        assert bc in ("rbc", "hc"), f"Boundary condition type {bc!r} not recognized!"
        self.bc = bc
        self.periodic = periodic
        self.scale = [aspect, 1.0]
        self.nu = get_nu(ra, pr, self.scale[1] * 2.0)
        self.ka = get_ka(ra, pr, self.scale[1] * 2.0)
        self.ra, self.pr, self.dt, self.time = ra, pr, dt, 0.0
        if periodic:  # navier.rs:336-428
            bx = lambda: fourier_r2c(nx)
            self.velx = Field2(Space2(bx(), cheb_dirichlet(ny)))
            self.vely = Field2(Space2(bx(), cheb_dirichlet(ny)))
            self.temp = Field2(Space2(bx(), cheb_dirichlet(ny) if bc == "rbc" else cheb_dirichlet_neumann(ny)))
            self.tempbc = bc_rbc(bx(), ny) if bc == "rbc" else bc_hc(bx(), ny, True)
            self.pres = Field2(Space2(bx(), chebyshev(ny)))
            self.pseu = Field2(Space2(bx(), cheb_neumann(ny)))
            self.field = Field2(Space2(bx(), chebyshev(ny)))
        else:  # navier.rs:215-308
            self.velx = Field2(Space2(cheb_dirichlet(nx), cheb_dirichlet(ny)))
            self.vely = Field2(Space2(cheb_dirichlet(nx), cheb_dirichlet(ny)))
            self.temp = Field2(Space2(cheb_neumann(nx), cheb_dirichlet(ny) if bc == "rbc" else cheb_dirichlet_neumann(ny)))
            self.tempbc = bc_rbc(chebyshev(nx), ny) if bc == "rbc" else bc_hc(chebyshev(nx), ny)
            self.pres = Field2(Space2(chebyshev(nx), chebyshev(ny)))
            self.pseu = Field2(Space2(cheb_neumann(nx), cheb_neumann(ny)))
            self.field = Field2(Space2(chebyshev(nx), chebyshev(ny)))
        for f in (self.velx, self.vely, self.temp, self.pres):
            f.scale(self.scale)
        sc = self.scale
        self.solver_hholtz = [
            HholtzAdi(self.velx, [dt * self.nu / sc[0] ** 2, dt * self.nu / sc[1] ** 2]),
            HholtzAdi(self.vely, [dt * self.nu / sc[0] ** 2, dt * self.nu / sc[1] ** 2]),
            HholtzAdi(self.temp, [dt * self.ka / sc[0] ** 2, dt * self.ka / sc[1] ** 2]),
        ]
        self.solver_pres = Poisson(self.pseu, [1.0 / sc[0] ** 2, 1.0 / sc[1] ** 2], eig=pois_eig)
        # rhs buffer shape: navier.rs:277 (confined) / :397 (periodic)
        self.rhs_shape = self.field.vhat.shape if periodic else self.temp.v.shape

    # -- initial conditions ----------------------------------------------------
    def _unit_coords(self, field):
        x, y = field.x
        return (x - x[0]) / (x[-1] - x[0]), (y - y[0]) / (y[-1] - y[0])

    def set_velocity(self, amp, m, n):
        """navier.rs:156-159 + functions.rs:85-125."""
        x, y = self._unit_coords(self.velx)
        self.velx.v = amp * np.outer(np.sin(np.pi * m * x), np.cos(np.pi * n * y))
        self.velx.forward()
        x, y = self._unit_coords(self.vely)
        self.vely.v = -amp * np.outer(np.cos(np.pi * m * x), np.sin(np.pi * n * y))
        self.vely.forward()

    def set_temperature(self, amp, m, n):
        """navier.rs:164-166."""
        x, y = self._unit_coords(self.temp)
        self.temp.v = -amp * np.outer(np.cos(np.pi * m * x), np.sin(np.pi * n * y))
        self.temp.forward()

    def init_random(self, amp, seeds=(1, 2, 3)):
        """navier.rs:171-182: U(-amp, amp) physical fields then forward (synthetic:
        numpy default_rng(seed) replaces ndarray-rand)."""
        for f, s in zip((self.temp, self.velx, self.vely), seeds):
            f.v = np.random.default_rng(s).uniform(-amp, amp, size=f.v.shape)
            f.forward()

    # -- equations (src/navier_stokes/navier_eq.rs) -----------------------------
    def _conv_term(self, u, field, deriv):
        """functions.rs:56-69:  u * backward(gradient)."""
        return u * self.field.space.backward(field.gradient(deriv, self.scale))

    def _conv(self, field, ux, uy, with_bc=False):
        """navier_eq.rs:60-101."""
        conv = self._conv_term(ux, field, [1, 0])
        conv = conv + self._conv_term(uy, field, [0, 1])
        if with_bc:
            conv = conv + self._conv_term(ux, self.tempbc, [1, 0])
            conv = conv + self._conv_term(uy, self.tempbc, [0, 1])
        self.field.v = conv
        self.field.forward()
        dealias(self.field.vhat)
        return np.array(self.field.vhat, copy=True)

    def div(self):
        """navier_eq.rs:19-24."""
        return self.velx.gradient([1, 0], self.scale) + self.vely.gradient([0, 1], self.scale)

    def div_norm(self):
        """navier_eq.rs:32-49 + functions.rs:24-35."""
        d = self.div()
        return float(np.sqrt(np.sum(d.real ** 2 + d.imag ** 2)))

    def update(self):
        """navier.rs:438-466."""
        dt = self.dt
        that = self.temp.to_ortho() + self.tempbc.to_ortho()
        self.velx.backward()
        self.vely.backward()
        ux, uy = self.velx.v.copy(), self.vely.v.copy()
        # solve_velx, navier_eq.rs:176-187
        rhs = self.velx.to_ortho()
        rhs = rhs - self.pres.gradient([1, 0], self.scale) * dt
        rhs = rhs - self._conv(self.velx, ux, uy) * dt
        self.velx.vhat = self.solver_hholtz[0].solve(rhs)
        # solve_vely, navier_eq.rs:190-203
        rhs = self.vely.to_ortho()
        rhs = rhs - self.pres.gradient([0, 1], self.scale) * dt
        rhs = rhs + that * dt
        rhs = rhs - self._conv(self.vely, ux, uy) * dt
        self.vely.vhat = self.solver_hholtz[1].solve(rhs)
        # projection, navier.rs:455-458
        div = self.div()
        self.pseu.vhat = self.solver_pres.solve(div)  # navier_eq.rs:158-162
        self.pseu.vhat[0, 0] = 0.0
        # correct_velocity(1.0), navier_eq.rs:117-125
        dp_dx = self.pseu.gradient([1, 0], self.scale) * (-1.0)
        dp_dy = self.pseu.gradient([0, 1], self.scale) * (-1.0)
        self.velx.vhat = self.velx.vhat + self.velx.space.from_ortho(dp_dx)
        self.vely.vhat = self.vely.vhat + self.vely.space.from_ortho(dp_dy)
        # update_pres, navier_eq.rs:137-143
        self.pres.vhat = self.pres.vhat + div * (-self.nu) + self.pseu.to_ortho() * (1.0 / dt)
        # solve_temp, navier_eq.rs:209-224
        rhs = self.temp.to_ortho()
        rhs = rhs + self.tempbc.gradient([2, 0], self.scale) * dt * self.ka
        rhs = rhs + self.tempbc.gradient([0, 2], self.scale) * dt * self.ka
        rhs = rhs - self._conv(self.temp, ux, uy, with_bc=True) * dt
        self.temp.vhat = self.solver_hholtz[2].solve(rhs)
        self.time += dt

    # -- diagnostics (src/navier_stokes/functions.rs:146-233, src/field/average.rs:26-62) -------
    @staticmethod
    def average_axis(field, axis):
        """``FieldBase::average_axis``: dx-weighted mean along ``axis`` (src/field/average.rs:26-35)."""
        length = abs(field.x[axis][-1] - field.x[axis][0])
        w = field.dx[axis] / length
        return np.tensordot(w, field.v, axes=(0, axis))

    @classmethod
    def average(cls, field):
        """``FieldBase::average`` (src/field/average.rs:53-59)."""
        length = abs(field.x[1][-1] - field.x[1][0])
        return float(np.sum(cls.average_axis(field, 0) * field.dx[1] / length))

    def eval_nu(self):
        """Nusselt number from the heat flux at the plates (functions.rs:146-168)."""
        f = self.field
        f.vhat = self.temp.to_ortho() + self.tempbc.to_ortho()
        f.vhat = f.gradient([0, 1], None) * (-2.0 / self.scale[1])
        f.backward()
        x_avg = self.average_axis(f, 0)
        return float((x_avg[-1] + x_avg[0]) / 2.0)

    def eval_nuvol(self):
        """Volumetric Nusselt number (functions.rs:175-207)."""
        f = self.field
        f.vhat = self.temp.to_ortho() + self.tempbc.to_ortho()
        f.backward()
        self.vely.backward()
        vely_temp = f.v * self.vely.v
        f.vhat = f.gradient([0, 1], None) / (-self.scale[1])
        f.backward()
        f.v = (f.v + vely_temp / self.ka) * 2.0 * self.scale[1]
        return self.average(f)

    def eval_re(self):
        """Reynolds number from the kinetic energy (functions.rs:215-233)."""
        self.velx.backward()
        self.vely.backward()
        f = self.field
        f.v = np.sqrt(self.velx.v ** 2 + self.vely.v ** 2) * (2.0 * self.scale[1] / self.nu)
        return self.average(f)

    def state(self):
        return {k: np.array(getattr(self, k).vhat, copy=True) for k in ("temp", "velx", "vely", "pres")}


# ---------------------------------------------------------------------------
# Multi-rank: slab ("pencil") decomposition (funspace Decomp2d)
# ---------------------------------------------------------------------------
def split_bounds(n, nprocs):
    """2decomp-style contiguous split: n // P each, remainder one extra to the
    HIGHEST ranks (unpinned by reference tests).  Returns inclusive
    (st, en) per rank as funspace ``Decomp2d.{x,y}_pencil.{st,en}``."""
    base, rem = divmod(n, nprocs)
    sizes = [base + (1 if r >= nprocs - rem else 0) for r in range(nprocs)]
    st = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    return [(int(s), int(s + z - 1)) for s, z in zip(st, sizes)]


class Decomp2d:
    """x-pencil: axis 0 complete, axis 1 split; y-pencil: axis 1 complete, axis 0
    split (src/field_mpi.rs:130-134, src/solver_mpi/poisson.rs:171)."""

    def __init__(self, shape, nprocs, rank):
        self.shape, self.nprocs, self.rank = tuple(shape), nprocs, rank
        self.b0 = split_bounds(shape[0], nprocs)
        self.b1 = split_bounds(shape[1], nprocs)

    def x_pencil(self, rank=None):
        st, en = self.b1[self.rank if rank is None else rank]
        return (slice(None), slice(st, en + 1))

    def y_pencil(self, rank=None):
        st, en = self.b0[self.rank if rank is None else rank]
        return (slice(st, en + 1), slice(None))


def dealias_mpi_xpen(vhat_x_pen, shape_spectral, dcp: Decomp2d, reference_quirk=False):
    """src/navier_stokes_mpi/functions.rs:75-97 on an x-pencil slab.  With
    ``reference_quirk`` the reference's strict ``n_y < en[1]`` test is kept (a rank
    whose last owned column equals n_y is then left unzeroed); the default is the
    serial rule, so that gather(local) == serial result."""
    n_x = shape_spectral[0] * 2 // 3
    n_y = shape_spectral[1] * 2 // 3
    vhat_x_pen[n_x:, :] = 0
    st, en = dcp.b1[dcp.rank]
    if (n_y < en) if reference_quirk else (n_y <= en):
        yst = n_y - st if n_y > st else 0
        vhat_x_pen[:, yst:] = 0
