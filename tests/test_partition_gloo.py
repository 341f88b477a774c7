"""The multi-GPU host logic on CPU: two gloo ranks exercise the slab partition, the halo
exchange and the CFL all-reduce of pyclaw_b200.parallel (PetClaw's DMDA replacement)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, periodic, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import petclaw as pyclaw
        from pyclaw_b200.parallel import slab_range
        mx, my, mbc, meqn = 7, 11, 2, 3
        x = pyclaw.Dimension('x', 0., 1., mx)
        y = pyclaw.Dimension('y', 0., 1., my)
        grid = pyclaw.Grid([x, y])
        state = pyclaw.State(grid, meqn, device='cpu')
        part = state._partition
        j0, j1 = slab_range(my, rank, world)
        assert (grid.y.nstart, grid.y.nend) == (j0, j1) and grid.y.ng == j1 - j0 and grid.x.ng == mx
        assert abs(grid.y.lowerg - j0 / my) < 1e-15
        assert len(grid.y.center) == j1 - j0 and abs(grid.y.center[0] - (j0 + 0.5) / my) < 1e-15
        state.set_mbc(mbc)
        # global field f(m,i,j) known in closed form
        f = lambda m, i, j: 1000. * m + 10. * i + 0.01 * j
        I, J = np.meshgrid(np.arange(mx), np.arange(j0, j1), indexing='ij')
        for m in range(meqn):
            state.q[m, :, :] = f(m, I, J)
        part.exchange(state._q, meqn, periodic=[False, periodic])
        qbc = np.asarray(state._q.padded())
        ok = True
        for g in range(mbc):
            jl = j0 - mbc + g            # global row of lower ghost g
            ju = j1 + g
            for (jg, jj) in ((jl, g), (ju, mbc + (j1 - j0) + g)):
                if 0 <= jg < my or periodic:
                    jw = jg % my
                    exp = np.stack([f(m, np.arange(mx), jw) for m in range(meqn)])
                    ok &= np.array_equal(qbc[:, mbc:-mbc, jj], exp)
                else:
                    ok &= np.all(qbc[:, mbc:-mbc, jj] == 0.0)     # untouched: physical BC's job
        cfl = torch.tensor([0.1 * (rank + 1)] + [0.0] * 15, dtype=torch.float64)
        part.allreduce_max(cfl)
        ok &= float(cfl[0]) == 0.1 * world
        full = part.gather_interior(state)
        ok &= full.shape == (meqn, mx, my) and full[2, 3, 9] == f(2, 3, 9)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [False, True])
def test_halo_exchange_two_ranks(periodic):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, periodic, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_slab_range_is_a_partition():
    from pyclaw_b200.parallel import slab_range
    for n in (8, 11, 8192, 8193):
        for size in (1, 2, 3, 8):
            r = [slab_range(n, k, size) for k in range(size)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(size - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _io_worker(rank, world, port, path, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import petclaw as pyclaw
        mx, my, meqn, maux = 6, 9, 2, 1
        grid = pyclaw.Grid([pyclaw.Dimension('x', 0., 1., mx), pyclaw.Dimension('y', 0., 2., my)])
        state = pyclaw.State(grid, meqn, maux, device='cpu')
        j0, j1 = grid.y.nstart, grid.y.nend
        f = lambda m, i, j: 100. * m + 10. * i + 0.125 * j
        I, J = np.meshgrid(np.arange(mx), np.arange(j0, j1), indexing='ij')
        for m in range(meqn):
            state.q[m, :, :] = f(m, I, J)
        state.aux[0, :, :] = f(7, I, J)
        state.t = 1.5
        sol = pyclaw.Solution(state)
        ok = pyclaw.Controller().output_format == 'petsc'
        for fmt in ('petsc', 'ascii'):
            sol.write(2, path, fmt, write_aux=True)
            dist.barrier()
            # every rank reads the frame back onto its own slab
            back = pyclaw.Solution(2, path=path, format=fmt, read_aux=True)
            ok &= isinstance(back.state, pyclaw.State) and back.state.grid.y.nstart == j0
            ok &= np.array_equal(np.asarray(back.q), np.asarray(state.q))
            ok &= np.array_equal(np.asarray(back.aux), np.asarray(state.aux))
            ok &= back.t == 1.5
            dist.barrier()
        if rank == 0:
            # the files hold the global field, written once
            import pyclaw as serial
            glob = serial.Solution(2, path=path, format='petsc', options={'state_class': serial.State})
            ok &= np.asarray(glob.q).shape == (meqn, mx, my) and float(np.asarray(glob.q)[1, 4, 8]) == f(1, 4, 8)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_frames_written_and_restarted_on_two_ranks(tmp_path):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_io_worker, args=(world, port, str(tmp_path), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _nd_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import petclaw as pyclaw
        ok = True
        mbc = 2
        # 1-D: the only dimension is partitioned; 3-D: z-slabs, halo "rows" are padded planes
        for dims in ([('x', 13)], [('x', 5), ('y', 4), ('z', 9)]):
            grid = pyclaw.Grid([pyclaw.Dimension(n, 0., 1., m) for n, m in dims])
            state = pyclaw.State(grid, 2, device='cpu')
            state.set_mbc(mbc)
            last = grid.dimensions[-1]
            k0, k1 = last.nstart, last.nend
            glob = np.arange(2 * np.prod([m for _, m in dims]), dtype=float).reshape([2] + [m for _, m in dims])
            state.q[...] = glob[..., k0:k1]
            state._partition.exchange(state._q, 2, periodic=[False] * (len(dims) - 1) + [True])
            qbc = np.asarray(state._q.padded())
            nloc = k1 - k0
            n = dims[-1][1]
            inner = (slice(None),) + (slice(mbc, -mbc),) * (len(dims) - 1)
            for g in range(mbc):
                lo = (k0 - mbc + g) % n
                hi = (k1 + g) % n
                ok &= np.array_equal(qbc[inner + (g,)], glob[..., lo])
                ok &= np.array_equal(qbc[inner + (mbc + nloc + g,)], glob[..., hi])
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_halo_exchange_in_1d_and_3d():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_nd_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
