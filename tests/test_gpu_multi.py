"""Multi-GPU parity (-m gpu, needs >= 2 GPUs): slab-decomposed Navier2D with peer-store transposes over
NVLink, one process per GPU, gathered state against the serial oracle (tests/dist_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("nx,ny,steps,periodic,mode,port", [
    (129, 129, 5, 0, 1, 29711), (128, 129, 5, 1, 1, 29712), (129, 129, 2, 0, 0, 29713), (257, 129, 3, 0, 1, 29714)])
def test_navier_slabs_match_serial_oracle(nx, ny, steps, periodic, mode, port):
    world = min(ngpus(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (a single rank is what tests/test_gpu_parity.py covers)")
    env = dict(os.environ, B2_TEST_EMU="0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")] + [str(a) for a in (nx, ny, steps, periodic, mode)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    assert r.stdout.count("worst_rel_err") == world
