"""Worker for the multi-rank tests: one process per rank (torch.distributed), slab-decomposed
Navier2D steps, gathered state compared with the serial oracle on every rank.

  CPU (tests/test_dist_gloo.py):   backend gloo, library = SIMT-emulator build, "GPUs" = processes
  GPU (tests/test_gpu_multi.py):   backend nccl/gloo, library = CUDA build, one GPU per rank
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    use_emu = os.environ.get("B2_TEST_EMU", "0") == "1"
    if use_emu:
        from tests import emu

        emu.activate()
    import numpy as np
    import torch
    import torch.distributed as dist

    import rustpde_mpi_b200 as b2
    from oracle import rustpde_oracle as o

    dist.init_process_group(backend="gloo" if use_emu else "cpu:gloo,cuda:nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    device = 0 if use_emu else int(os.environ.get("LOCAL_RANK", rank))
    if not use_emu:
        torch.cuda.set_device(device)
    nx, ny, steps, periodic, mode = (int(v) for v in sys.argv[1:6])
    ctx = b2.Context.distributed(device, heap_bytes=(200 * (nx + 16) * (ny + 16) * 8) // world + (8 << 20))
    nav = b2.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=bool(periodic), ctx=ctx)
    nav.set_mode(mode)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    nav.update(steps)
    got = nav.gather_state()
    dn = nav.div_norm()
    eig = None if periodic else b2.poisson_eig(b2.CHEB_NEUMANN, nx, 1.0)
    ref = o.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=bool(periodic), pois_eig=eig)
    ref.set_velocity(0.2, 1.0, 1.0)
    ref.set_temperature(0.2, 1.0, 1.0)
    for _ in range(steps):
        ref.update()
    worst = 0.0
    for k, v in ref.state().items():
        assert got[k].shape == v.shape, (k, got[k].shape, v.shape)
        err = float(np.linalg.norm((got[k] - v).ravel()) / np.linalg.norm(v.ravel()))
        worst = max(worst, err)
    assert abs(dn - ref.div_norm()) <= 1e-8 * max(1.0, ref.div_norm()), (dn, ref.div_norm())
    # callback() diagnostics on the slabs (src/field_mpi/average.rs:15-61: partial dx-weighted sums + all_gather_sum)
    for name in ("eval_nu", "eval_nuvol", "eval_re"):
        a, b = getattr(nav, name)(), getattr(ref, name)()
        assert abs(a - b) <= 1e-10 * abs(b), (name, a, b)
    print(f"rank {rank}/{world}: nx={nx} ny={ny} steps={steps} periodic={periodic} mode={mode} worst_rel_err={worst:.3e}", flush=True)
    assert worst < 1e-10, worst
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
