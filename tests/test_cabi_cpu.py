"""CPU-side tests of the boundary: the C-ABI library builds, loads and exports every symbol that
include/b200pde.h declares (no compute calls: there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

import rustpde_mpi_b200
from rustpde_mpi_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200pde.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_header_symbol():
    path = build.build()
    assert os.path.exists(path)
    so = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(so, s), f"{s} declared in include/b200pde.h but not exported"


def test_python_binding_table_matches_header():
    assert sorted(_lib.SYMBOLS) == header_symbols()


def test_version_and_error_string():
    l = _lib.lib()
    assert l.b2_version() >= 1
    assert isinstance(l.b2_last_error(), bytes)


def test_host_only_entry_point_poisson_matrices():
    """b2_host_poisson_matrices is pure host code: check it against the oracle's matrices
    (src/field.rs:195-249 + src/solver/poisson.rs:65-74)."""
    import numpy as np

    from oracle import rustpde_oracle as o

    n = 33
    m = n - 2
    a0 = np.zeros((m, m)); c0 = np.zeros((m, m))
    st = _lib.lib().b2_host_poisson_matrices(2, n, 0.7, a0.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                             c0.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert st == 0
    f = o.Field2(o.Space2(o.cheb_neumann(n), o.cheb_neumann(n)))
    mat_a, mat_b, _, _ = f.ingredients_for_poisson(0)
    np.testing.assert_allclose(c0, mat_a, rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(a0, mat_b * 0.7, rtol=1e-14, atol=1e-16)


def test_poisson_eig_parity_split_is_equivalent():
    import numpy as np

    lam, fwd, bwd = rustpde_mpi_b200.poisson_eig(2, 65, 1.0, parity_split=False)
    lam2, fwd2, bwd2 = rustpde_mpi_b200.poisson_eig(2, 65, 1.0, parity_split=True)
    np.testing.assert_allclose(lam, lam2, rtol=1e-9, atol=1e-9)
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal((63, 4))
    x1 = bwd @ ((fwd @ rhs) / (lam[:, None] - 3.0))
    x2 = bwd2 @ ((fwd2 @ rhs) / (lam2[:, None] - 3.0))
    np.testing.assert_allclose(x1, x2, rtol=1e-8, atol=1e-10)


def test_no_device_fails_loudly():
    """Without a CUDA device the product path must raise, never fall back."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(rustpde_mpi_b200.B2Error):
        rustpde_mpi_b200.Context(0)


def test_rust_sys_bindings_cover_the_header():
    """rust/b200pde-sys/src/lib.rs is generated from the header (tools/gen_rust_sys.py; no Rust toolchain in this image): every
    symbol is declared, with the header's argument count, and the committed file is what the generator produces."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    path = os.path.join(ROOT, "rust", "b200pde-sys", "src", "lib.rs")
    text = open(path).read()
    decl = dict(re.findall(r"pub fn (b2_[a-z0-9_]+)\((.*?)\) ->", text))
    assert sorted(decl) == header_symbols()
    for ret, name, args in gen.prototypes():
        n_c = 0 if args.strip() == "void" else len(args.split(","))
        n_r = 0 if not decl[name].strip() else len(decl[name].split(","))
        assert n_c == n_r, name
    before = text
    gen.main()
    assert open(path).read() == before, "rust/b200pde-sys/src/lib.rs is stale: run tools/gen_rust_sys.py"


def test_cpp_driver_compiles_against_the_header():
    """examples/cpp_driver/navier_rbc.cpp: plain C++ over the C ABI (no CUDA headers); it must compile and link here."""
    import subprocess
    import tempfile

    build.build()
    out = os.path.join(tempfile.mkdtemp(), "navier_rbc")
    libdir = os.path.join(ROOT, "rustpde_mpi_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cpp_driver", "navier_rbc.cpp"),
                    "-o", out, "-L", libdir, "-lb200pde", "-ldl", f"-Wl,-rpath,{libdir}"], check=True)
    assert os.path.exists(out)
