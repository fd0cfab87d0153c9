"""Variant sweep on one GPU: the same Navier2D config under several tuning environments in ONE process (the host LAPACK
setup is done once).  Prints ms/step, GEMM ms, lane ms and the per-op cycle breakdown for every variant.

  python tools/sweep.py C4 "name1:K=V,K=V" "name2:..."      (a bare "base" = no overrides)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustpde_mpi_b200 as b2  # noqa: E402
from bench import CONFIGS  # noqa: E402

cfg = sys.argv[1]
variants = sys.argv[2:] or ["base"]
nx, ny, ra, dt, per = CONFIGS[cfg]
steps = int(os.environ.get("SWEEP_STEPS", "10"))
prof = os.environ.get("SWEEP_OPPROF", "1") != "0"
ctx = b2.Context(0)
eig = None if per else b2.poisson_eig(b2.CHEB_NEUMANN, nx, 1.0)
TUNING = [k for k in os.environ if k.startswith("B2_") and k != "B2_EIG_CACHE"]
for v in variants:
    name, _, kv = v.partition(":")
    for k in list(os.environ):
        if k.startswith("B2_") and k != "B2_EIG_CACHE" and k not in TUNING:
            del os.environ[k]
    for item in filter(None, kv.split(",")):
        k, _, val = item.partition("=")
        os.environ[k] = val
    try:
        nav = b2.Navier2D(nx, ny, ra, 1.0, dt, 1.0, "rbc", periodic=per, ctx=ctx, pois_eig=eig)
        nav.set_mode(1)
        nav.update(3)
        ctx.sync()
        ctx.timer_start(); nav.update(steps); ms = ctx.timer_stop() / steps
        ctx.profile(True); nav.update(3); gemm = ctx.profile(False) / 3
        line = f"{cfg} {name:28s} {ms:8.3f} ms/step  gemm {gemm:6.3f}  lane {ms - gemm:7.3f}  div {nav.div_norm():.3e}"
        if prof:
            nav.set_mode(3); nav.update(1)
            ctx.opprof(True); nav.update(2); p = ctx.opprof(False)
            line += "  | " + " ".join(f"{k}={c / n / 1e3:.1f}k" for k, (c, n) in sorted(p.items(), key=lambda kv: -kv[1][0]))
        print(line, flush=True)
        nav.close()
    except Exception as e:  # noqa: BLE001
        print(f"{cfg} {name:28s} FAILED: {e!r}", flush=True)
