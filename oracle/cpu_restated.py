"""Loader for oracle/cpu_restated.cpp (the C++/OpenMP restatement of the reference's update(); TEST INFRASTRUCTURE
and bench.py's CPU baseline only -- the product never imports this).  The library is built into oracle/_cpu/."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_restated.cpp")
OUT = os.path.join(HERE, "_cpu", "librustpde_cpu.so")
FIELDS = {"temp": 0, "velx": 1, "vely": 2, "pres": 3, "pseu": 4, "tempbc": 5}


def build(force=False):
    """g++ -O3 -fopenmp; x86-64-v3 (AVX2 + FMA) so that the file built in one container runs on another host."""
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O3", "-std=c++17", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-ldl"], check=True)
    return OUT


def openblas_path():
    """The OpenBLAS (ILP64) inside the numpy wheel -- the reference links OpenBLAS through ndarray-linalg."""
    hits = glob.glob(os.path.join(os.path.dirname(os.path.dirname(np.__file__)), "numpy.libs", "libscipy_openblas64_*.so"))
    return hits[0] if hits else ""


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.rc_navier_create.restype = C.c_void_p
        L.rc_navier_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rc_navier_div_norm.restype = C.c_double
        L.rc_last_error.restype = C.c_char_p
        for f in (L.rc_navier_destroy, L.rc_navier_update, L.rc_navier_set_v, L.rc_navier_vhat_shape, L.rc_navier_get_vhat,
                  L.rc_navier_set_vhat, L.rc_navier_div_norm):
            f.argtypes = None
        _lib = L
    return _lib


class Navier2D:
    """Same surface as oracle.rustpde_oracle.Navier2D for the parts bench.py and the tests use."""

    def __init__(self, nx, ny, ra, pr, dt, aspect, bc="rbc", periodic=False, pois_eig=None, threads=0):
        assert bc == "rbc"
        L = lib()
        self.blas = bool(L.rc_init(openblas_path().encode(), int(threads)))
        self.threads = L.rc_threads()
        self.nx, self.ny, self.periodic = nx, ny, periodic
        if not periodic:
            if pois_eig is None or isinstance(pois_eig, str):
                from . import rustpde_oracle as o

                f = o.Field2(o.Space2(o.cheb_neumann(nx), o.cheb_neumann(ny)))
                mass, lap, _, _ = f.ingredients_for_poisson(0)
                lam, fwd, bwd = o.parity_eig(lap * (1.0 / aspect ** 2), mass)
                if abs(lam[0]) < 1e-10:
                    lam = lam - 1e-10
                pois_eig = (lam, fwd, bwd)
            self._eig = [np.ascontiguousarray(a, dtype=np.float64) for a in pois_eig]
            ptrs = [a.ctypes.data_as(C.c_void_p) for a in self._eig]
        else:
            ptrs = [None, None, None]
        self._h = C.c_void_p(L.rc_navier_create(nx, ny, ra, pr, dt, aspect, int(periodic), *ptrs))
        if not self._h:
            raise RuntimeError(L.rc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rc_navier_destroy(self._h)
            self._h = None

    def set_v(self, name, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.nx, self.ny)
        lib().rc_navier_set_v(self._h, FIELDS[name], v.ctypes.data_as(C.c_void_p))

    def _unit(self):
        """Unit coordinates of navier.rs:156-166 (functions.rs:85-125): (x - x0) / (x_last - x0) per axis."""
        x = 2.0 * np.pi * np.arange(self.nx) / self.nx if self.periodic else -np.cos(np.pi * np.arange(self.nx) / (self.nx - 1))
        y = -np.cos(np.pi * np.arange(self.ny) / (self.ny - 1))
        return (x - x[0]) / (x[-1] - x[0]), (y - y[0]) / (y[-1] - y[0])

    def set_velocity(self, amp, m, n):
        x, y = self._unit()
        self.set_v("velx", amp * np.outer(np.sin(np.pi * m * x), np.cos(np.pi * n * y)))
        self.set_v("vely", -amp * np.outer(np.cos(np.pi * m * x), np.sin(np.pi * n * y)))

    def set_temperature(self, amp, m, n):
        x, y = self._unit()
        self.set_v("temp", -amp * np.outer(np.cos(np.pi * m * x), np.sin(np.pi * n * y)))

    def init_random(self, amp, seeds=(1, 2, 3)):
        for name, s in zip(("temp", "velx", "vely"), seeds):
            self.set_v(name, np.random.default_rng(s).uniform(-amp, amp, size=(self.nx, self.ny)))

    def vhat(self, name):
        r, c, cx = C.c_int(), C.c_int(), C.c_int()
        lib().rc_navier_vhat_shape(self._h, FIELDS[name], C.byref(r), C.byref(c), C.byref(cx))
        out = np.empty((r.value, c.value), dtype=np.complex128 if cx.value else np.float64)
        lib().rc_navier_get_vhat(self._h, FIELDS[name], out.ctypes.data_as(C.c_void_p))
        return out

    def set_vhat(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.complex128 if self.periodic else np.float64)
        assert a.shape == self.vhat(name).shape
        lib().rc_navier_set_vhat(self._h, FIELDS[name], a.ctypes.data_as(C.c_void_p))

    def update(self, steps=1):
        lib().rc_navier_update(self._h, int(steps))

    def div_norm(self):
        return float(lib().rc_navier_div_norm(self._h))

    OPS = ("backward", "forward", "to_ortho", "from_ortho", "gradient_10", "gradient_02", "hholtz_adi", "poisson")

    def time_ops(self, calls=3):
        """Seconds per call of the standalone operators on this problem's temperature / pseudo-pressure spaces (state untouched)."""
        sec = (C.c_double * 8)()
        lib().rc_navier_time_ops(self._h, int(calls), sec)
        return dict(zip(self.OPS, (float(v) for v in sec)))

    def state(self):
        return {k: self.vhat(k) for k in ("temp", "velx", "vely", "pres")}


if __name__ == "__main__":
    print(build(force=True))
