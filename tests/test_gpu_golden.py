"""
End-to-end parity through the PyClaw API (``import pyclaw`` -> pyclaw_b200): the
reference's own regression problems (test/test_examples.py) run on the GPU and are
compared with the reference's golden files and with the CPU oracle driver.
The scripts below follow test/acoustics/*/acoustics.py and test/euler/2d/shockbubble.py.
"""
import os

import numpy as np
import pytest
import torch

import problems
from oracle import pyclaw_oracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _acoustics1d(solver_type, **opts):
    import pyclaw
    if solver_type == 'classic':
        solver = pyclaw.ClawSolver1D()
    else:
        solver = pyclaw.SharpClawSolver1D()
        solver.weno_order = 5
    for k, v in opts.items():
        setattr(solver, k, v)
    x = pyclaw.Dimension('x', 0.0, 1.0, 100)
    grid = pyclaw.Grid(x)
    state = pyclaw.State(grid, 2)
    rho, bulk = 1.0, 1.0
    state.aux_global['rho'] = rho
    state.aux_global['bulk'] = bulk
    state.aux_global['zz'] = np.sqrt(rho * bulk)
    state.aux_global['cc'] = np.sqrt(rho / bulk)
    xc = grid.x.center
    state.q[0, :] = np.exp(-100 * (xc - 0.75) ** 2) * np.cos(0 * (xc - 0.75))
    state.q[1, :] = 0.
    solver.mwaves = 2
    solver.limiters = [4] * solver.mwaves
    solver.dt_initial = grid.d[0] / state.aux_global['cc'] * 0.1
    solver.bc_lower[0] = pyclaw.BC.periodic
    solver.bc_upper[0] = pyclaw.BC.periodic
    claw = pyclaw.Controller()
    claw.keep_copy = True
    claw.nout = 5
    claw.output_format = None
    claw.tfinal = 1.0
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.run()
    q0 = np.asarray(claw.frames[0].state.q).reshape([-1])
    qfinal = np.asarray(claw.frames[claw.nout].state.q).reshape([-1])
    return grid.d[0] * np.sum(np.abs(qfinal - q0)), claw


def test_acoustics1d_classic():
    err, claw = _acoustics1d('classic')
    assert abs(err - 0.00104856594174) < 1e-14          # test_examples.py:59-66 (tolerance 1e-5)


def test_acoustics1d_sharpclaw_weno17():
    # test_examples.py:162-169: 0.000163221216565 within the reference's own 1e-5
    err, claw = _acoustics1d('sharpclaw', weno_order=17)
    assert abs(err - 0.000163221216565) < 1e-5, err
    # and the oracle driver with the same tables gives the same frames bit for bit
    from pyclaw_b200.weno_tables import tables
    pb = problems.acoustics1d(100)
    s = po.OracleSolver("sharpclaw", 1, po.RP_ACOUSTICS, pb["params"], 2)
    s.bc_lower = s.bc_upper = [po.BC_PERIODIC]
    s.dt_initial = pb["dt_initial"]
    s.weno_order, s.weno_tables = 17, tables(9, 'f32')
    fr = s.run(pb["q"], None, pb["d"], 1.0, 5)
    assert np.array_equal(np.asarray(claw.frames[-1].state.q), fr[-1])


def test_acoustics1d_sharpclaw():
    err, _ = _acoustics1d('sharpclaw', weno_literals='f64')
    assert abs(err - 0.000298935748775) < 1e-12          # test_examples.py:125-150 (tolerance 1e-5)
    err, _ = _acoustics1d('sharpclaw')                   # REAL(4) literals, as gfortran compiles weno.f90
    assert abs(err - 0.000298935748775) < 1e-5


def _acoustics2d(solver_type, **opts):
    import pyclaw
    if solver_type == 'classic':
        solver = pyclaw.ClawSolver2D()
    else:
        solver = pyclaw.SharpClawSolver2D()
    solver.cfl_max = 0.5
    solver.cfl_desired = 0.45
    solver.mwaves = 2
    solver.dim_split = 1
    solver.limiters = [4] * solver.mwaves
    for k, v in opts.items():
        setattr(solver, k, v)
    solver.bc_lower[0] = pyclaw.BC.outflow
    solver.bc_upper[0] = pyclaw.BC.outflow
    solver.bc_lower[1] = pyclaw.BC.outflow
    solver.bc_upper[1] = pyclaw.BC.outflow
    mx = my = 100
    x = pyclaw.grid.Dimension('x', -1.0, 1.0, mx)
    y = pyclaw.grid.Dimension('y', -1.0, 1.0, my)
    grid = pyclaw.grid.Grid([x, y])
    state = pyclaw.State(grid, 3)
    rho, bulk = 1.0, 4.0
    cc = np.sqrt(bulk / rho)
    zz = rho * cc
    state.aux_global.update(rho=rho, bulk=bulk, zz=zz, cc=cc)
    Y, X = np.meshgrid(grid.y.center, grid.x.center)
    r = np.sqrt(X ** 2 + Y ** 2)
    width = 0.2
    state.q[0, :, :] = (np.abs(r - 0.5) <= width) * (1. + np.cos(np.pi * (r - 0.5) / width))
    state.q[1, :, :] = 0.
    state.q[2, :, :] = 0.
    solver.dt_initial = np.min(grid.d) / state.aux_global['cc'] * solver.cfl_desired
    claw = pyclaw.Controller()
    claw.keep_copy = True
    claw.output_format = None
    claw.tfinal = 0.12
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.nout = 10
    claw.run()
    return claw.frames[claw.nout].state.q[0, :, :]


def test_acoustics2d_classic_golden():
    pressure = _acoustics2d('classic')
    verify_x = np.loadtxt(os.path.join(GOLD, 'acoustics2D_solution'))
    assert np.linalg.norm(np.asarray(pressure) - verify_x) < 2e-14   # test_examples.py:239-254


def test_acoustics2d_sharpclaw_golden():
    verify_x = np.loadtxt(os.path.join(GOLD, 'ac_sc_solution'))
    p = _acoustics2d('sharpclaw')                                    # default: PyWENO, REAL(4) literals
    assert np.linalg.norm(np.asarray(p) - verify_x) < 1e-4           # test_examples.py:333-376
    p = _acoustics2d('sharpclaw', lim_type=3)                        # hand-written weno5
    assert np.linalg.norm(np.asarray(p) - verify_x) < 1e-12


def _shockbubble(**opts):
    import pyclaw
    from pyclaw.clawpack import ClawSolver2D
    gamma, gamma1 = problems.GAMMA, problems.GAMMA1
    rinf, vinf, einf = problems.shock_state()

    def shockbc(state, dim, t, qbc, mbc):
        if dim.nstart == 0:
            for i in range(mbc):
                qbc[0, i, ...] = rinf
                qbc[1, i, ...] = rinf * vinf
                qbc[2, i, ...] = 0.
                qbc[3, i, ...] = einf
                qbc[4, i, ...] = 0.

    def euler_rad_src(solver, state, dt):
        problems.euler_rad_src(torch, state.q, state.aux, dt)

    pb = problems.shockbubble()
    x = pyclaw.Dimension('x', 0.0, 2.0, 160)
    y = pyclaw.Dimension('y', 0.0, 0.5, 40)
    grid = pyclaw.Grid([x, y])
    state = pyclaw.State(grid, 5, 1)
    state.aux_global['gamma'] = gamma
    state.aux_global['gamma1'] = gamma1
    state.q[...] = pb["q"]
    state.aux[...] = pb["aux"]
    solver = ClawSolver2D()
    solver.cfl_max = 0.5
    solver.cfl_desired = 0.45
    solver.mwaves = 5
    solver.limiters = [4, 4, 4, 4, 2]
    solver.dt_initial = 0.005
    solver.user_bc_lower = shockbc
    solver.step_src = euler_rad_src
    solver.src_split = 1
    solver.bc_lower[0] = pyclaw.BC.custom
    solver.bc_upper[0] = pyclaw.BC.outflow
    solver.bc_lower[1] = pyclaw.BC.reflecting
    solver.bc_upper[1] = pyclaw.BC.outflow
    solver.aux_bc_lower[0] = pyclaw.BC.outflow
    solver.aux_bc_upper[0] = pyclaw.BC.outflow
    solver.aux_bc_lower[1] = pyclaw.BC.outflow
    solver.aux_bc_upper[1] = pyclaw.BC.outflow
    for k, v in opts.items():
        setattr(solver, k, v)
    claw = pyclaw.Controller()
    claw.keep_copy = True
    claw.output_format = None
    claw.tfinal = 0.2
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.nout = 1
    status = claw.run()
    return claw.frames[claw.nout].state.q, status


def test_shockbubble_golden_bit_exact():
    q, status = _shockbubble()
    verify_x = np.loadtxt(os.path.join(GOLD, 'sb_density'))
    assert status['numsteps'] == 170
    # test_examples.py:385-397 asks for max abs < 1e-12; the GPU path is bit-exact
    assert np.max(np.abs(np.asarray(q[0]) - verify_x)) == 0.0


def _oracle_shockbubble(dim_split, order_trans, src):
    pb = problems.shockbubble()
    s = po.OracleSolver("classic", 2, po.RP_EULER5, pb["params"], 5)
    s.cfl_max, s.cfl_desired = 0.5, 0.45
    s.limiters = pb["limiters"]
    s.dt_initial = pb["dt_initial"]
    s.dim_split, s.order_trans = dim_split, order_trans
    s.bc_lower = [po.BC_CUSTOM, po.BC_REFLECTING]
    s.bc_upper = [po.BC_OUTFLOW, po.BC_OUTFLOW]
    s.user_bc_lower = problems.shockbc_numpy
    if src:
        s.step_src = lambda solver, state, dt: problems.euler_rad_src(np, state["q"], state["aux"], dt)
    frames = s.run(pb["q"], pb["aux"], pb["d"], pb["tfinal"], pb["nout"])
    return frames[-1], s.total


@pytest.mark.parametrize("order_trans", [1, 2])
def test_shockbubble_unsplit_vs_oracle(order_trans):
    # the application's own setting (apps/euler/2d/shockbubble: dim_split=False, order_trans=2)
    q, status = _shockbubble(dim_split=False, order_trans=order_trans)
    qo, total = _oracle_shockbubble(False, order_trans, True)
    assert np.array_equal(np.asarray(q), qo)


def _shallow(kind, **opts):
    import pyclaw
    pb = problems.shallow2d(60, 60)
    solver = pyclaw.ClawSolver2D() if kind == 'classic' else pyclaw.SharpClawSolver2D()
    solver.mwaves = 3
    solver.limiters = pyclaw.limiters.tvd.MC
    solver.bc_lower[0] = pyclaw.BC.outflow
    solver.bc_upper[0] = pyclaw.BC.reflecting
    solver.bc_lower[1] = pyclaw.BC.outflow
    solver.bc_upper[1] = pyclaw.BC.reflecting
    solver.dim_split = 1
    for k, v in opts.items():
        setattr(solver, k, v)
    x = pyclaw.Dimension('x', -2.5, 2.5, 60)
    y = pyclaw.Dimension('y', -2.5, 2.5, 60)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3)
    state.aux_global['grav'] = 1.0
    state.q[...] = pb["q"]
    claw = pyclaw.Controller()
    claw.tfinal = 1.0
    claw.keep_copy = True
    claw.output_format = None
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.nout = 2
    claw.run()
    return np.asarray(claw.frames[-1].state.q)


def _oracle_shallow(kind, **opts):
    pb = problems.shallow2d(60, 60)
    s = po.OracleSolver(kind, 2, po.RP_SHALLOW, [1.0], 3)
    s.limiters = 4
    s.bc_lower = [po.BC_OUTFLOW, po.BC_OUTFLOW]
    s.bc_upper = [po.BC_REFLECTING, po.BC_REFLECTING]
    s.dim_split = True
    for k, v in opts.items():
        setattr(s, k, v)
    return s.run(pb["q"], None, pb["d"], 1.0, 2)[-1]


@pytest.mark.parametrize("opts", [dict(dim_split=1), dict(dim_split=0, order_trans=2),
                                  dict(dim_split=0, order_trans=1), dict(dim_split=0, order_trans=0, cfl_max=0.5, cfl_desired=0.45)])
def test_shallow_classic_vs_oracle(opts):
    assert np.array_equal(_shallow('classic', **opts), _oracle_shallow('classic', **opts))


@pytest.mark.parametrize("ti", ['SSP33', 'SSP104', 'Euler'])
def test_shallow_sharpclaw_vs_oracle(ti):
    # the reference's default cfl 2.45/2.5 is tuned for SSP104 in 1-D; use values that are
    # stable for each integrator on this 2-D problem so that the comparison is NaN free
    cfl = {'Euler': (0.5, 0.4), 'SSP33': (0.6, 0.5), 'SSP104': (1.3, 1.2)}[ti]
    o = dict(time_integrator=ti, cfl_max=cfl[0], cfl_desired=cfl[1])
    qg = _shallow('sharpclaw', **o)
    qo = _oracle_shallow('sharpclaw', **o)
    assert not np.isnan(qo).any()
    assert np.array_equal(qg, qo)


def test_rejected_step_is_rolled_back():
    # a far too large first dt must be rejected and retried (solver.py:686-698)
    import pyclaw
    q1 = _acoustics2d('classic')
    q2 = _acoustics2d('classic', dt_initial_scale=None) if False else None
    pb = problems.acoustics2d()
    s = po.OracleSolver("classic", 2, po.RP_ACOUSTICS, pb["params"], 2)
    s.cfl_max, s.cfl_desired = 0.5, 0.45
    s.limiters = [4, 4]
    s.bc_lower = [po.BC_OUTFLOW] * 2
    s.bc_upper = [po.BC_OUTFLOW] * 2
    s.dt_initial = 10 * pb["dt_initial"]
    fo = s.run(pb["q"], None, pb["d"], 0.06, 2)
    assert s.total["rejected"] >= 1

    solver = pyclaw.ClawSolver2D()
    solver.cfl_max, solver.cfl_desired, solver.mwaves, solver.dim_split = 0.5, 0.45, 2, 1
    solver.limiters = [4, 4]
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    x = pyclaw.Dimension('x', -1.0, 1.0, 100)
    y = pyclaw.Dimension('y', -1.0, 1.0, 100)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3)
    state.aux_global.update(rho=1.0, bulk=4.0, zz=2.0, cc=2.0)
    state.q[...] = pb["q"]
    solver.dt_initial = 10 * pb["dt_initial"]
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format, claw.tfinal, claw.nout = True, None, 0.06, 2
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.run()
    assert np.array_equal(np.asarray(claw.frames[-1].state.q), fo[-1])


def _sphere_api(mx=40, my=20, tfinal=10.0, nout=10):
    """test/shallow_sphere/shallow_4_Rossby_Haurwitz_wave.py on the GPU.  setaux.f / qinit.f
    are init-time helpers of the application (f2py `problem` module in the reference); here
    their restatement supplies the initial arrays.  src2.f (Coriolis, 4-stage RK + tangent
    plane projection) is the per-step source term and runs on the device."""
    import pyclaw
    pb = problems.sphere_problem(mx, my)
    full = pb["auxbc_full"]
    mbc = 2

    def qbc_lower_y(state, dim, t, qbc, mbc):
        for j in range(mbc):
            qbc[:, :, j] = torch.flip(qbc[:, :, 2 * mbc - 1 - j], dims=[1])

    def qbc_upper_y(state, dim, t, qbc, mbc):
        my_ = state.grid.ng[1]
        for j in range(mbc):
            qbc[:, :, my_ + mbc + j] = torch.flip(qbc[:, :, my_ + mbc - 1 - j], dims=[1])

    def auxbc_lower_y(state, dim, t, auxbc, mbc):
        auxbc[:, :, :mbc] = full[:, :, :mbc]

    def auxbc_upper_y(state, dim, t, auxbc, mbc):
        auxbc[:, :, -mbc:] = full[:, :, -mbc:]

    df = float(np.float32(12.600576))          # "12.600576e0" is a REAL(4) literal in src2.f:38
    six = None

    def src2(solver, state, dt):
        nonlocal six
        q, aux = state.q, state.aux
        er = [aux[13], aux[14], aux[15]]       # = mapc2p(cell centre), what src2.f recomputes
        if six is None:
            six = torch.full((), 6.0, dtype=q.dtype, device=q.device)

        def project():
            qn = er[0] * q[1] + er[1] * q[2] + er[2] * q[3]
            q[1] = q[1] - qn * er[0]
            q[2] = q[2] - qn * er[1]
            q[3] = q[3] - qn * er[2]
        project()
        fcor = df * er[2]
        RK = []
        hu, hv, hw = q[1], q[2], q[3]
        for st in range(4):
            if st > 0:
                hu = q[1] + 0.5 * RK[st - 1][0]
                hv = q[2] + 0.5 * RK[st - 1][1]
                hw = q[3] + 0.5 * RK[st - 1][2]
            RK.append((fcor * dt * (er[2] * hv - er[1] * hw),
                       dt * fcor * (er[0] * hw - er[2] * hu),
                       dt * fcor * (er[1] * hu - er[0] * hv)))
        for m in range(3):
            q[m + 1] = q[m + 1] + torch.div(RK[0][m] + 2.0 * RK[1][m] + 2.0 * RK[2][m] + RK[3][m], six)
        project()

    solver = pyclaw.ClawSolver2D()
    solver.rp = pyclaw.riemann.shallow_sphere
    solver.bc_lower[0] = pyclaw.BC.periodic
    solver.bc_upper[0] = pyclaw.BC.periodic
    solver.bc_lower[1] = pyclaw.BC.custom
    solver.bc_upper[1] = pyclaw.BC.custom
    solver.user_bc_lower = qbc_lower_y
    solver.user_bc_upper = qbc_upper_y
    solver.aux_bc_lower[0] = pyclaw.BC.periodic
    solver.aux_bc_upper[0] = pyclaw.BC.periodic
    solver.aux_bc_lower[1] = pyclaw.BC.custom
    solver.aux_bc_upper[1] = pyclaw.BC.custom
    solver.user_aux_bc_lower = auxbc_lower_y
    solver.user_aux_bc_upper = auxbc_upper_y
    solver.dim_split = 0
    solver.order_trans = 2
    solver.mwaves = 3
    solver.src_split = 2
    solver.step_src = src2
    solver.limiters = pyclaw.limiters.tvd.MC
    x = pyclaw.Dimension('x', -3.0, 1.0, mx)
    y = pyclaw.Dimension('y', -1.0, 1.0, my)
    state = pyclaw.State(pyclaw.Grid([x, y]), 4, 16)
    state.aux_global['g'] = problems.SPHERE_G
    state.aux[:, :, :] = pb["aux"]
    state.mcapa = 0
    state.q[:, :, :] = pb["q"]
    claw = pyclaw.Controller()
    claw.keep_copy = True
    claw.output_format = None
    claw.outstyle = 1
    claw.nout = nout
    claw.tfinal = tfinal
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.run()
    return np.asarray(claw.frames[claw.nout].state.q)


def test_shallow_sphere_golden():
    # test_examples.py:456-472 asks for a Frobenius norm < 1e-4 against test/swsphere_height
    q = _sphere_api()
    gold = np.loadtxt(os.path.join(GOLD, 'swsphere_height'))
    assert np.linalg.norm(q[0] - gold) < 1e-15


def test_shallow_sphere_vs_oracle():
    import test_oracle_golden as tog
    frames, s = tog._oracle_sphere(tfinal=2.0)
    q = _sphere_api(tfinal=2.0)
    assert np.array_equal(q, frames[-1])


def test_cuda_graph_replay_is_bit_identical():
    """The captured-and-replayed step (clawpack.py `_hyperbolic_sequence`) against eager launches."""
    import pyclaw
    outs = []
    for use_graph in (True, False):
        saved = pyclaw.ClawSolver2D.use_cuda_graph
        pyclaw.ClawSolver2D.use_cuda_graph = use_graph
        try:
            outs.append(np.asarray(_acoustics2d('classic', dim_split=0, order_trans=2)))
            outs.append(np.asarray(_acoustics2d('classic')))
        finally:
            pyclaw.ClawSolver2D.use_cuda_graph = saved
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])
    # and the graphs were really used
    solver = pyclaw.ClawSolver2D()
    solver.mwaves = 2
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    x = pyclaw.Dimension('x', -1.0, 1.0, 64)
    y = pyclaw.Dimension('y', -1.0, 1.0, 64)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3)
    state.aux_global.update(rho=1.0, bulk=4.0, zz=2.0, cc=2.0)
    state.q[...] = problems.acoustics2d(64, 64)["q"]
    sol = pyclaw.Solution(state)
    solver.setup(sol)
    solver.dt = 0.001
    for _ in range(8):
        solver.evolve_to_time(sol)
    assert 1 <= len(solver._graphs) <= 6


def test_shallow_sphere_app_module():
    """pyclaw_b200.apps.shallow_sphere: vectorised numpy setaux/qinit agree with the restated
    Fortran helpers to round-off, and the packaged set-up reproduces the golden height."""
    import pyclaw
    from pyclaw_b200.apps import shallow_sphere as app
    pb = problems.sphere_problem(40, 20)
    dx, dy = pb["d"]
    aux = app.setaux(2, 40, 20, -3.0, -1.0, dx, dy)
    assert np.abs(aux[1:] - pb["auxbc_full"][1:]).max() < 1e-13
    # kappa is a spherical excess computed through acos(beta ~ 1): ill-conditioned, last-bit
    # differences between libm and numpy are amplified to ~1e-9
    assert np.abs(aux[0] - pb["auxbc_full"][0]).max() < 1e-7
    q = app.qinit(40, 20, -3.0, -1.0, dx, dy)
    assert np.abs(q - pb["q"]).max() < 1e-14
    state, solver = app.setup(pyclaw)
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format, claw.nout, claw.tfinal = True, None, 10, 10
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.run()
    gold = np.loadtxt(os.path.join(GOLD, 'swsphere_height'))
    # libm vs numpy transcendental functions differ in the last bits of aux: 1e-4 is the
    # reference's own tolerance (test_examples.py:466), we are many orders below it
    assert np.linalg.norm(np.asarray(claw.frames[-1].state.q[0]) - gold) < 1e-8


def test_sphere_src2_kernel_vs_tensor_ops_vs_oracle():
    import pyclaw
    from pyclaw_b200.apps import shallow_sphere as app
    state, solver = app.setup(pyclaw, 64, 32)
    sol = pyclaw.Solution(state)
    solver.setup(sol)
    q0 = np.asfortranarray(np.asarray(state.q))
    aux = np.asfortranarray(np.asarray(state.aux))
    dt = 0.0123
    app.src2(solver, state, dt)
    a = np.asarray(state.q).copy()
    state.q[...] = q0
    app.src2_torch(solver, state, dt)
    b = np.asarray(state.q).copy()
    qo = q0.copy("F")
    po.sphere_src2(qo, aux, -3.0, -1.0, state.grid.d[0], state.grid.d[1], dt)
    assert np.array_equal(a, b)
    # the oracle recomputes the radial vector with libm-free mapc2p: identical to aux(14:16)
    # only if aux came from the same mapc2p; here aux is numpy's -> compare to round-off
    assert np.abs(a - qo).max() < 1e-15


@pytest.mark.parametrize("solver_type,ic", [('classic', '2-shock'), ('classic', 'dam-break'), ('sharpclaw', 'dam-break')])
def test_shallow1d_app_vs_oracle(solver_type, ic):
    """apps/shallow/1d/shallow1D.py, line by line, against the oracle driver (the reference has no
    golden for it; the 1-D Roe solver is external: parity unpinned, GPU vs oracle bit for bit)."""
    import pyclaw
    solver = pyclaw.ClawSolver1D() if solver_type == 'classic' else pyclaw.SharpClawSolver1D()
    solver.mwaves = 2
    solver.limiters = pyclaw.limiters.tvd.vanleer
    solver.kernel_language = 'Fortran'
    solver.bc_lower[0] = pyclaw.BC.outflow
    solver.bc_upper[0] = pyclaw.BC.outflow
    if solver_type == 'sharpclaw':
        solver.cfl_max, solver.cfl_desired = 1.3, 1.2
    mx = 500
    grid = pyclaw.Grid(pyclaw.Dimension('x', -5.0, 5.0, mx))
    state = pyclaw.State(grid, 2)
    state.aux_global['grav'] = 1.0
    xc = grid.x.center
    hl, ul, hr, ur = (3., 0., 1., 0.) if ic == 'dam-break' else (1., 1., 1., -1.)
    q0 = np.zeros((2, mx), order='F')
    q0[0] = hl * (xc <= 0.) + hr * (xc > 0.)
    q0[1] = hl * ul * (xc <= 0.) + hr * ur * (xc > 0.)
    state.q[...] = q0
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format, claw.tfinal = True, None, 2.0
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.run()
    qg = np.asarray(claw.frames[-1].q)

    s = po.OracleSolver(solver_type, 1, po.RP_SHALLOW, [1.0], 2)
    s.bc_lower = s.bc_upper = [po.BC_OUTFLOW]
    if solver_type == 'classic':
        s.limiters = 3
    else:
        s.cfl_max, s.cfl_desired = 1.3, 1.2
    qo = s.run(q0, None, [grid.d[0]], 2.0, 10)[-1]
    assert not np.isnan(qo).any()
    assert np.array_equal(qg, qo)
    if ic == 'dam-break':       # exact middle state of the 3:1 dam break with g = 1
        assert abs(qg[0, mx // 2] - 1.848576) < 2e-3


def test_baseline_config0_acoustics1d_800_cells():
    """BASELINE.json configs[0]: 1-D acoustics, ClawSolver1D, MC limiter, 800 cells -- the
    reference's CPU-runnable case; GPU through the pyclaw API vs the oracle driver, bit for bit."""
    import pyclaw
    mx = 800
    solver = pyclaw.ClawSolver1D()
    solver.mwaves, solver.limiters = 2, [4, 4]
    solver.bc_lower[0] = solver.bc_upper[0] = pyclaw.BC.periodic
    grid = pyclaw.Grid(pyclaw.Dimension('x', 0.0, 1.0, mx))
    state = pyclaw.State(grid, 2)
    state.aux_global.update(rho=1.0, bulk=1.0, zz=1.0, cc=1.0)
    xc = grid.x.center
    q0 = np.zeros((2, mx), order='F')
    q0[0] = np.exp(-100 * (xc - 0.75) ** 2)
    state.q[...] = q0
    solver.dt_initial = grid.d[0] * 0.1
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format, claw.tfinal, claw.nout = True, None, 1.0, 5
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.run()
    s = po.OracleSolver("classic", 1, po.RP_ACOUSTICS, [1.0, 1.0, 1.0, 1.0], 2)
    s.limiters = [4, 4]
    s.bc_lower = s.bc_upper = [po.BC_PERIODIC]
    s.dt_initial = grid.d[0] * 0.1
    fr = s.run(q0, None, [grid.d[0]], 1.0, 5)
    for k in range(6):
        assert np.array_equal(np.asarray(claw.frames[k].q), fr[k]), k
    # one period: the pulse is back where it started, second-order accurate
    assert grid.d[0] * np.abs(fr[-1] - fr[0]).sum() < 2e-4


def test_two_weno_orders_alive_at_once():
    """The WENO tables are one set per process; interleaving solvers of different orders must
    still give each its own (sharpclaw.py re-uploads when another upload came in between)."""
    import pyclaw

    def make(order):
        solver = pyclaw.SharpClawSolver1D()
        solver.weno_order = order
        solver.mwaves = 2
        solver.bc_lower[0] = solver.bc_upper[0] = pyclaw.BC.periodic
        grid = pyclaw.Grid(pyclaw.Dimension('x', 0.0, 1.0, 100))
        state = pyclaw.State(grid, 2)
        state.aux_global.update(rho=1.0, bulk=1.0, zz=1.0, cc=1.0)
        state.q[0, :] = np.exp(-100 * (grid.x.center - 0.75) ** 2)
        state.q[1, :] = 0.
        solver.dt_initial = 0.001
        sol = pyclaw.Solution(state)
        solver.setup(sol)
        solver.dt = solver.dt_initial
        return solver, sol

    a, sa = make(9)
    b, sb = make(13)
    for _ in range(5):                      # interleaved
        a.evolve_to_time(sa)
        b.evolve_to_time(sb)
    a2, sa2 = make(9)
    for _ in range(5):                      # alone
        a2.evolve_to_time(sa2)
    assert np.array_equal(np.asarray(sa.q), np.asarray(sa2.q))
    assert not np.array_equal(np.asarray(sa.q), np.asarray(sb.q))


def test_acoustics2d_sharpclaw_weno7_vs_oracle():
    """test/acoustics/2d/homogeneous with SharpClaw and weno_order = 7 (mbc = 4), SSP104."""
    from pyclaw_b200.weno_tables import tables
    p = np.asarray(_acoustics2d('sharpclaw', weno_order=7))
    pb = problems.acoustics2d()
    s = po.OracleSolver("sharpclaw", 2, po.RP_ACOUSTICS, pb["params"], 2)
    s.cfl_max, s.cfl_desired = 0.5, 0.45
    s.bc_lower = [po.BC_OUTFLOW] * 2
    s.bc_upper = [po.BC_OUTFLOW] * 2
    s.dt_initial = pb["dt_initial"]
    s.weno_order, s.weno_tables = 7, tables(4, 'f32')
    qo = s.run(pb["q"], None, pb["d"], 0.12, 10)[-1]
    assert not np.isnan(qo).any()
    assert np.array_equal(p, qo[0])
    # close to, but not the same as, the fifth-order result
    p5 = np.asarray(_acoustics2d('sharpclaw'))
    assert 1e-8 < np.abs(p - p5).max() < 5e-2


@pytest.mark.parametrize("lim", [1, 4])
def test_shallow_sharpclaw_tvd2_vs_oracle(lim):
    """lim_type = 1 (second-order TVD reconstruction) through the API, SSP33, against the oracle."""
    o = dict(time_integrator='SSP33', cfl_max=0.6, cfl_desired=0.5, lim_type=1, limiters=[lim] * 3)
    qg = _shallow('sharpclaw', **o)
    qo = _oracle_shallow('sharpclaw', **o)
    assert np.isfinite(qo).all()
    assert np.array_equal(qg, qo)


@pytest.mark.parametrize("ti", ['SSP33', 'SSP104'])
def test_acoustics1d_sharpclaw_wave_based_vs_oracle(ti):
    """char_decomp = 1 through the API (1-D acoustics, the reference's test problem) against the
    oracle; the wave-based scheme converges like the component-wise one on this smooth problem."""
    # the reference's default cfl 2.45 / 2.5 is SSP104's; SSP33 needs a smaller one to be stable
    cfl = {'SSP33': (0.6, 0.5), 'SSP104': (2.5, 2.45)}[ti]
    err, claw = _acoustics1d('sharpclaw', char_decomp=1, time_integrator=ti, cfl_max=cfl[0], cfl_desired=cfl[1])
    pb = problems.acoustics1d(100)
    s = po.OracleSolver("sharpclaw", 1, po.RP_ACOUSTICS, pb["params"], 2)
    s.char_decomp, s.time_integrator = 1, ti
    s.cfl_max, s.cfl_desired = cfl
    s.bc_lower = s.bc_upper = [po.BC_PERIODIC]
    s.dt_initial = pb["dt_initial"]
    frames = s.run(pb["q"], None, pb["d"], pb["tfinal"], pb["nout"])
    assert np.array_equal(np.asarray(claw.frames[-1].q), frames[-1])
    assert err < 4e-4   # component-wise WENO5 gives 2.99e-4 on this problem, the wave-based form 2.82e-4
