"""CPU SIMT emulator of the CUDA sources (test infrastructure only)."""
import os


def activate():
    """Point the ctypes loader of rustpde_mpi_b200 at the emulator build.  Only tests call this;
    the package itself has no switch and no CPU path."""
    from . import build_emu
    import rustpde_mpi_b200._lib as L

    path = build_emu.build()
    assert L._lib is None or L.LIB_PATH == path, "library already loaded"
    L.LIB_PATH = path
    return path
