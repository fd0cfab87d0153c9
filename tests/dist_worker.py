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
    bc = "hc" if "hc" in sys.argv[6:] else "rbc"
    ctx = b2.Context.distributed(device, heap_bytes=(200 * (nx + 16) * (ny + 16) * 8) // world + (8 << 20))
    nav = b2.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, bc, periodic=bool(periodic), ctx=ctx)
    nav.set_mode(mode)
    nav.set_velocity(0.2, 1.0, 1.0)
    nav.set_temperature(0.2, 1.0, 1.0)
    nav.update(steps)
    got = nav.gather_state()
    dn = nav.div_norm()
    eig = None if periodic else b2.poisson_eig(b2.CHEB_NEUMANN, nx, 1.0)
    ref = o.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, bc, periodic=bool(periodic), pois_eig=eig)
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
    if "extras" in sys.argv[6:]:
        # HholtzMpi / PoissonMpi (src/solver_mpi/{hholtz,poisson}.rs): the field solvers on slabs, input and output as local rows
        for name, kinds in (("Hholtz", (1, 1)), ("Poisson", (2, 2))):
            k0 = 4 if periodic else kinds[0]
            fo = o.Field2(o.Space2(o.Base(k0, nx), o.Base(kinds[1], ny)))
            fg = b2.Field2(b2.Space2((k0, nx), (kinds[1], ny), ctx=ctx))
            c = [0.37, 1.3] if name == "Hholtz" else [1.0, 1.0]
            e = None
            if not periodic:
                e = b2.hholtz_eig(k0, nx, c[0]) if name == "Hholtz" else b2.poisson_eig(k0, nx, c[0])
            so_, sg = getattr(o, name)(fo, c, eig=e), getattr(b2, name)(fg, c)
            sh = fo.space.to_ortho(fo.vhat).shape
            rng = np.random.default_rng(11)
            rhs = rng.standard_normal(sh) + (1j * rng.standard_normal(sh) if periodic else 0)
            inp = b2.DeviceArray(fg.space, b2.ORTHO)
            r0, cnt = inp.local_rows()
            inp.set(rhs[r0:r0 + cnt])
            x = ctx.all_gather_rows(sg.solve(inp).get())
            xo = so_.solve(rhs)
            if name == "Poisson":
                x[0, 0] = 0; xo[0, 0] = 0
            err = float(np.abs(x - xo).max() / np.abs(xo).max())
            assert err < 1e-10, (name, err)
        # gather / scatter of the slabs (src/field_mpi.rs:363-453): scatter from rank 0, transform on the slabs, gather back
        fo = o.Field2(o.Space2(o.Base(4 if periodic else 1, nx), o.Base(2, ny)))
        fg = b2.Field2(b2.Space2((4 if periodic else 1, nx), (2, ny), ctx=ctx))
        vg = np.random.default_rng(5).standard_normal((nx, ny))
        fg.scatter_physical_root(vg if rank == 0 else None, root=0)
        assert fg.nrank() == rank and fg.nprocs() == world
        assert np.array_equal(fg.all_gather_physical(), vg)
        assert np.array_equal(fg.get_coords_local(0), fg.x[0][fg.local_slice(b2.PHYSICAL)])
        fg.forward()
        fo.v = vg; fo.forward()
        got = fg.gather_spectral_root(root=world - 1)
        assert (got is None) == (rank != world - 1)
        if got is not None:
            assert float(np.abs(got - fo.vhat).max() / np.abs(fo.vhat).max()) < 1e-10
        fg.scatter_spectral_root(fo.vhat if rank == 0 else None, root=0)
        assert np.array_equal(fg.all_gather_spectral(), fo.vhat)
        # average_axis / average on slabs (src/field_mpi/average.rs:15-61): partial sums added over the ranks (axis 0), row parts
        # concatenated (axis 1)
        fg.scatter_physical_root(vg if rank == 0 else None, root=0)
        fo.v = vg
        for ax in (0, 1):
            a, b = fg.average_axis(ax), o.Navier2D.average_axis(fo, ax)
            assert a.shape == b.shape and float(np.abs(a - b).max()) < 1e-13, (ax, a.shape, b.shape)
        assert abs(fg.average() - o.Navier2D.average(fo)) < 1e-13
        # FourierC2c x ChebDirichlet on slabs: complex physical rows, forward / gradient / backward against the serial oracle
        if not periodic:
            fo = o.Field2(o.Space2(o.fourier_c2c(48), o.cheb_dirichlet(ny)))
            fg = b2.Field2(b2.Space2(b2.fourier_c2c(48), b2.cheb_dirichlet(ny), ctx=ctx))
            rng = np.random.default_rng(9)
            vg = rng.standard_normal((48, ny)) + 1j * rng.standard_normal((48, ny))
            fg.v = vg[fg.local_slice(b2.PHYSICAL)]
            fg.forward()
            fo.v = vg; fo.forward()
            got = fg.all_gather_spectral()
            assert float(np.abs(got - fo.vhat).max() / np.abs(fo.vhat).max()) < 1e-10
            gr = ctx.all_gather_rows(fg.gradient([1, 1]).get())
            ref = fo.gradient([1, 1])
            assert float(np.abs(gr - ref).max() / np.abs(ref).max()) < 1e-10
            fg.backward(); fo.backward()   # (not the identity: the composite base projects onto the boundary conditions)
            assert float(np.abs(fg.all_gather_physical() - fo.v).max() / np.abs(fo.v).max()) < 1e-10
        # snapshot / restart on slabs (src/field_mpi/io.rs): gathered write on rank 0, every rank reads its rows back
        import tempfile
        fn = os.path.join(tempfile.gettempdir(), f"b2_snap_{os.environ.get('MASTER_PORT', '0')}.npz")
        nav.write(fn)
        nav2 = b2.Navier2D(nx, ny, 1e5, 1.0, 0.01, 1.0, "rbc", periodic=bool(periodic), ctx=ctx)
        nav2.read(fn)
        for k, v in nav.state().items():
            assert np.array_equal(nav2.state()[k], v), k
        assert abs(nav2.get_time() - nav.get_time()) < 1e-15
        # the collective allocator notices ranks that release arrays at different moments (last: the heaps differ afterwards)
        sp = b2.Space2((1, nx), (1, ny), ctx=ctx)
        keep = [b2.DeviceArray(sp, b2.ORTHO) for _ in range(2)]
        if rank == 1:
            keep[0].close()
        try:
            b2.DeviceArray(sp, b2.ORTHO)
            raise SystemExit("expected the symmetric-heap check to fail")
        except b2.B2Error as e:
            assert "symmetric heap diverged" in str(e), e
        print(f"rank {rank}/{world}: extras ok (HholtzMpi, PoissonMpi, gather / scatter, c2c, snapshot, heap check)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
