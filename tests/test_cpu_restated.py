"""The C++/OpenMP restatement of update() (oracle/cpu_restated.cpp, bench.py's CPU baseline and the fast checker of the
large GPU parity tests) against the numpy oracle, which is pinned to the reference's golden vectors."""
import numpy as np
import pytest

from oracle import cpu_restated as cr
from oracle import rustpde_oracle as o


def _pair(nx, ny, ra, dt, periodic, threads=0):
    ref = o.Navier2D(nx, ny, ra, 1.0, dt, 1.0, "rbc", periodic=periodic)
    eig = None
    if not periodic:
        t = ref.solver_pres.solver
        eig = (t.lam[0], t.fwd[0], t.bwd[0])
    cpp = cr.Navier2D(nx, ny, ra, 1.0, dt, 1.0, "rbc", periodic=periodic, pois_eig=eig, threads=threads)
    ref.init_random(0.1)
    cpp.init_random(0.1)
    return ref, cpp


def _worst(ref, cpp):
    worst = 0.0
    for k, v in ref.state().items():
        worst = max(worst, float(np.abs(cpp.vhat(k) - v).max() / max(np.abs(v).max(), 1e-300)))
    return worst


@pytest.mark.parametrize("case", [(129, 129, 1e5, 1e-2, False), (128, 129, 1e5, 1e-2, True), (65, 33, 1e4, 1e-2, False)],
                         ids=["confined129", "periodic128x129", "confined65x33"])
def test_update_matches_numpy_oracle(case):
    ref, cpp = _pair(*case)
    assert _worst(ref, cpp) < 1e-12   # forward transforms of the initial fields
    for _ in range(3):
        ref.update()
    cpp.update(3)
    assert _worst(ref, cpp) < 1e-11
    assert abs(cpp.div_norm() - ref.div_norm()) < 1e-10 * max(1.0, ref.div_norm())


def test_single_thread_equals_all_threads():
    _, a = _pair(65, 65, 1e5, 1e-2, False, threads=1)
    assert a.threads == 1
    a.update(2)
    sa = a.state()
    _, b = _pair(65, 65, 1e5, 1e-2, False, threads=0)
    b.update(2)
    for k, v in b.state().items():
        assert np.array_equal(v, sa[k])   # lanes are independent: the thread count must not change a single bit


def test_time_ops_leaves_the_state_alone():
    """bench.py's CPU ms / transform, ms / solve run on copies: the run's state must not move."""
    from oracle import cpu_restated as cr

    nav = cr.Navier2D(65, 65, 1e5, 1.0, 0.01, 1.0, "rbc", threads=2)
    nav.set_velocity(0.2, 1.0, 1.0); nav.set_temperature(0.2, 1.0, 1.0)
    nav.update(2)
    before = nav.state()
    sec = nav.time_ops(2)
    assert set(sec) == set(cr.Navier2D.OPS) and all(v > 0 for v in sec.values())
    after = nav.state()
    assert all(np.array_equal(before[k], after[k]) for k in before)
    navp = cr.Navier2D(64, 65, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=True, threads=2)
    navp.set_velocity(0.2, 1.0, 1.0); navp.set_temperature(0.2, 1.0, 1.0)
    navp.update(1)
    assert all(v > 0 for v in navp.time_ops(1).values())
