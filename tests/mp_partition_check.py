"""Run under torchrun with N ranks (one GPU each): the slab-partitioned run must equal the
single-GPU run BIT FOR BIT (the reference compares its 6-rank run with the serial golden
file at 1e-14, test/test_examples.py:264-277)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems  # noqa: E402


def run(pc, kind, q0, aux0, mx, my, opts):
    rinf, vinf, einf = problems.shock_state()

    def shockbc(state, dim, t, qbc, mbc):
        if dim.nstart == 0:
            qbc[0, :mbc] = rinf
            qbc[1, :mbc] = rinf * vinf
            qbc[2, :mbc] = 0.
            qbc[3, :mbc] = einf
            qbc[4, :mbc] = 0.

    x = pc.Dimension('x', 0.0, 2.0, mx)
    y = pc.Dimension('y', 0.0, 0.5, my)
    grid = pc.Grid([x, y])
    if kind == 'euler':
        state = pc.State(grid, 5)
        state.aux_global.update(gamma=problems.GAMMA, gamma1=problems.GAMMA1)
        solver = pc.ClawSolver2D()
        solver.mwaves = 5
        solver.limiters = [4, 4, 4, 4, 2]
        solver.cfl_max, solver.cfl_desired = 0.5, 0.45
        solver.dt_initial = 0.005
        solver.user_bc_lower = shockbc
        solver.bc_lower[0] = pc.BC.custom
        solver.bc_upper[0] = pc.BC.outflow
        solver.bc_lower[1] = pc.BC.reflecting
        solver.bc_upper[1] = pc.BC.outflow
    else:
        state = pc.State(grid, 3)
        state.aux_global['grav'] = 1.0
        solver = pc.SharpClawSolver2D()
        solver.mwaves = 3
        solver.cfl_max, solver.cfl_desired = 0.6, 0.5
        solver.bc_lower[0] = pc.BC.outflow
        solver.bc_upper[0] = pc.BC.reflecting
        solver.bc_lower[1] = pc.BC.periodic
        solver.bc_upper[1] = pc.BC.periodic
        solver.dt_initial = 0.001
    for k, v in opts.items():
        setattr(solver, k, v)
    j0, j1 = grid.y.nstart, grid.y.nend
    state.q[...] = q0[:, :, j0:j1]
    claw = pc.Controller()
    claw.output_format = None
    claw.tfinal = 0.05
    claw.nout = 2
    claw.solution = pc.Solution(state)
    claw.solver = solver
    status = claw.run()
    part = state._partition
    q = part.gather_interior(state) if part is not None else np.asarray(state.q)
    return np.asarray(q), status['numsteps']


def run_sphere(pc, mx, my):
    """BASELINE config 5 on a slab partition: periodic x, pole-fold custom y BCs on the edge
    ranks, 16 aux components with capa, Strang-split fused src2."""
    from pyclaw_b200.apps import shallow_sphere as app
    state, solver = app.setup(pc, mx, my)
    solver.dt_initial = 0.1 * state.grid.d[0] / 4.0
    claw = pc.Controller()
    claw.output_format = None
    claw.tfinal = 0.02
    claw.nout = 2
    claw.solution = pc.Solution(state)
    claw.solver = solver
    status = claw.run()
    part = state._partition
    q = part.gather_interior(state) if part is not None else np.asarray(state.q)
    return np.asarray(q), status['numsteps']


def main():
    import petclaw
    rank, world = petclaw.init('nccl')
    import pyclaw
    mx, my = 160, 64
    pb = problems.shockbubble(mx, my)
    qs = problems.smooth_state("shallow", (mx, my), seed=5)
    cases = [('euler', pb["q"], dict(dim_split=False, order_trans=2)),
             ('euler', pb["q"], dict(dim_split=True)),
             ('shallow', qs, dict(time_integrator='SSP33')),
             ('shallow', qs, dict(time_integrator='SSP104', cfl_max=1.3, cfl_desired=1.2))]
    ok = True
    for kind, q0, opts in cases:
        qp, nsteps = run(petclaw, kind, q0, None, mx, my, opts)
        if rank == 0:
            qser, nser = run(pyclaw, kind, q0, None, mx, my, opts)
            same = np.array_equal(qp, qser) and nsteps == nser and not np.isnan(qser).any()
            print("case %s %s: %d ranks, %d steps, bit-identical=%s maxdiff=%g"
                  % (kind, opts, world, nsteps, same, np.abs(qp - qser).max()), flush=True)
            ok &= same
    qp, nsteps = run_sphere(petclaw, 64, 32)
    if rank == 0:
        qser, nser = run_sphere(pyclaw, 64, 32)
        same = np.array_equal(qp, qser) and nsteps == nser and not np.isnan(qser).any() and nser >= 3
        print("case sphere: %d ranks, %d steps, bit-identical=%s maxdiff=%g"
              % (world, nsteps, same, np.abs(qp - qser).max()), flush=True)
        ok &= same
    # the reference's applications (examples/apps.py) on the slab partition
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import apps
    for name, kw in (("shallow1d", dict(mx=200, tfinal=0.5)),
                     ("stegoton", dict(layers=30, tfinal=5.0, solver_type='sharpclaw')),
                     ("annulus", dict(mx=20, my=60, tfinal=0.2)),
                     ("vc_acoustics2d", dict(mx=48, my=40, tfinal=0.1)),
                     ("psystem", dict(cells_per_layer=8, tfinal=0.15)),
                     ("acoustics3d", dict(mx=32, my=4, mz=8 * world, tfinal=0.3)),
                     # single-pass unsplit kernel under the row-range calls of the overlapped step
                     ("acoustics2d", dict(mx=64, my=72, dim_split=False, tfinal=0.08)),
                     # unsplit 3-D (step3 + flux3 + rpt3 / rptt3), z-slabs
                     ("acoustics3d", dict(test='het', mx=14, my=12, mz=6 * world, tfinal=0.3, upper_bc=1))):
        cp = apps.APPS[name](petsc=True, **kw)
        st = cp.frames[-1].state
        qp = np.asarray(st._partition.gather_interior(st))
        if rank == 0:
            cs = apps.APPS[name](petsc=False, **kw)
            qs = np.asarray(cs.frames[-1].q)
            same = np.array_equal(qp, qs) and np.isfinite(qs).all()
            print("app %s: %d ranks, bit-identical=%s maxdiff=%g" % (name, world, same, np.abs(qp - qs).max()), flush=True)
            ok &= same
    flag = torch.tensor([1.0 if ok else 0.0], device='cuda')
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
