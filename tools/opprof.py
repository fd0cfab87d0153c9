"""Per-op cycle breakdown of the lane kernel for one Navier2D config (thread-0 clock64 deltas per op)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustpde_mpi_b200 as b2
from bench import CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
nx, ny, ra, dt, per = CONFIGS[cfg]
ctx = b2.Context(0)
nav = b2.Navier2D(nx, ny, ra, 1.0, dt, 1.0, "rbc", periodic=per, ctx=ctx)
nav.init_random(0.1)
nav.set_mode(3)
nav.update(3)
ctx.opprof(True)
steps = 3
ctx.timer_start(); nav.update(steps); ms = ctx.timer_stop()
prof = ctx.opprof(False)
tot = sum(c for c, _ in prof.values())
print(f"{cfg}: {ms/steps:.3f} ms/step (no graph, profiling on)")
for k, (c, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:10s} cycles/call={c/n:10.0f} calls/step={n/steps:10.0f} share={c/tot:6.1%}")
