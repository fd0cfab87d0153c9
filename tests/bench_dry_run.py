"""Dry run of bench.py's main path on the CPU (test infrastructure): the library is the SIMT-emulator build, the CUDA-only
pieces of torch are stubbed, the event timer is the wall clock.  Checks the control flow and the JSON contract of the repo arm
-- it measures nothing.  Usage: python tests/bench_dry_run.py <config> [bench.py flags ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import emu  # noqa: E402

emu.activate()
import torch  # noqa: E402

torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
import torch.distributed as _dist  # noqa: E402

_init = _dist.init_process_group
_dist.init_process_group = lambda backend=None, **k: _init(backend="gloo", **k)   # no NCCL without a GPU

import bench  # noqa: E402
import rustpde_mpi_b200 as b2  # noqa: E402

_t = {}


def _timer_start(self):
    self.sync()
    _t["t"] = time.perf_counter()


def _timer_stop(self):
    self.sync()
    return (time.perf_counter() - _t["t"]) * 1e3


_profile = b2.Context.profile


def _profile_ms(self, on):
    r = _profile(self, on)
    return r if on else 1.0   # the emulator has no events: a non-zero GEMM time keeps the arithmetic of the line finite


b2.Context.timer_start, b2.Context.timer_stop, b2.Context.profile = _timer_start, _timer_stop, _profile_ms
bench.CONFIGS["T0"] = (65, 65, 1e5, 1e-2, False)     # emulator-sized stand-ins for the confined / periodic workloads
bench.CONFIGS["T0p"] = (64, 65, 1e5, 1e-2, True)
bench.CONFIGS["T1"] = (129, 129, 1e5, 1e-2, False)   # smallest confined size with a thread layout on 8 ranks
sys.argv = ["bench.py", "--config", sys.argv[1], *sys.argv[2:]]
bench.main()
