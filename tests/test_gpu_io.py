"""Controller frame output and restart on the GPU."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(pyclaw, fixed_dt=True):
    solver = pyclaw.ClawSolver2D()
    solver.mwaves = 2
    solver.dim_split = 0
    solver.order_trans = 2
    solver.limiters = [4] * solver.mwaves
    for i in range(2):
        solver.bc_lower[i] = pyclaw.BC.outflow
        solver.bc_upper[i] = pyclaw.BC.reflecting
    mx, my = 60, 44
    grid = pyclaw.Grid([pyclaw.Dimension('x', -1.0, 1.0, mx), pyclaw.Dimension('y', -1.0, 1.0, my)])
    state = pyclaw.State(grid, 3)
    rho, bulk = 1.0, 4.0
    cc = np.sqrt(bulk / rho)
    state.aux_global.update(rho=rho, bulk=bulk, zz=rho * cc, cc=cc)
    Y, X = np.meshgrid(grid.y.center, grid.x.center)
    r = np.sqrt(X ** 2 + Y ** 2)
    state.q[0, :, :] = (np.abs(r - 0.5) <= 0.2) * (1. + np.cos(np.pi * (r - 0.5) / 0.2))
    state.q[1, :, :] = 0.
    state.q[2, :, :] = 0.
    solver.dt_initial = np.min(grid.d) / cc * 0.4
    if fixed_dt:
        # a power of two, so that output times and the clipped last step of every output
        # interval (solver.py:655) are exact and a restarted run sees the same dt sequence
        solver.dt_initial = 2.0 ** -8
        solver.dt_variable = False
    return solver, state


def test_controller_writes_frames_and_restarts_bit_exact(tmp_path):
    import pyclaw
    solver, state = _setup(pyclaw)
    dt = solver.dt_initial
    claw = pyclaw.Controller()
    claw.keep_copy = True
    claw.output_format = ['ascii', 'petsc']
    claw.outdir = str(tmp_path / '_output')
    claw.tfinal = 12 * dt
    claw.nout = 3
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.run()
    for k in range(4):
        for name in ('fort.t%04d', 'fort.q%04d', 'claw.pkl%04d', 'claw.ptc%04d'):
            assert os.path.exists(os.path.join(claw.outdir, name % k)), name % k
        ref = np.asarray(claw.frames[k].q)
        a = pyclaw.Solution(k, path=claw.outdir, format='ascii')
        b = pyclaw.Solution(k, path=claw.outdir, format='petsc')
        assert np.array_equal(np.asarray(b.q), ref)
        np.testing.assert_allclose(np.asarray(a.q), ref, rtol=5e-9, atol=1e-300)
        assert abs(a.t - claw.frames[k].t) < 1e-8 * max(1.0, abs(a.t)) and b.t == claw.frames[k].t
    # restart from frame 1 (binary format keeps every bit) and run the remaining frames
    solver2, _ = _setup(pyclaw)
    claw2 = pyclaw.Controller()
    claw2.keep_copy = True
    claw2.output_format = None
    claw2.solution = pyclaw.Solution(1, path=claw.outdir, format='petsc')
    assert claw2.solution.aux_global['bulk'] == 4.0
    claw2.solver = solver2
    claw2.start_frame = 1
    claw2.tfinal = 12 * dt
    claw2.nout = 2
    claw2.run()
    assert np.array_equal(np.asarray(claw2.frames[-1].q), np.asarray(claw.frames[-1].q))
    # refusing to overwrite
    claw.overwrite = False
    with pytest.raises(Exception):
        claw.run()


def test_compute_p_frames_and_functionals(tmp_path):
    import pyclaw
    solver, state = _setup(pyclaw, fixed_dt=False)
    state.mp = 1
    state.mF = 1

    def compute_p(st):
        st.p[0, :, :] = st.q[0, :, :] * 2.0

    def compute_F(st):
        st.F[0, :, :] = st.q[0, :, :] * st.grid.d[0] * st.grid.d[1]

    claw = pyclaw.Controller()
    claw.output_format = 'ascii'
    claw.outdir = str(tmp_path / '_output')
    claw.outdir_p = str(tmp_path / '_output' / '_p')
    claw.F_path = str(tmp_path / '_output' / 'F.txt')
    claw.compute_p, claw.compute_F = compute_p, compute_F
    claw.tfinal, claw.nout = 0.05, 2
    claw.keep_copy = True
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.run()
    p = pyclaw.Solution(2, path=claw.outdir_p, file_prefix='claw_p')
    np.testing.assert_allclose(np.asarray(p.q)[0], 2.0 * np.asarray(claw.frames[2].q)[0], rtol=5e-9, atol=1e-300)
    F = np.loadtxt(claw.F_path)
    assert F.shape == (3, 2)
    # F.txt rows are "t sum|F_0|" (controller.py:307-317, state.py sum_F)
    dxdy = state.grid.d[0] * state.grid.d[1]
    for k in range(3):
        assert abs(F[k, 0] - claw.frames[k].t) < 1e-15
        expect = np.abs(np.asarray(claw.frames[k].q)[0] * dxdy).sum()
        assert abs(F[k, 1] - expect) < 1e-13


def test_gauges_record_every_step(tmp_path):
    """grid.add_gauges / solver.write_gauge_values (grid.py:519-545, solver.py:731-741): one line
    ``t p...`` per accepted step plus the initial one, values from compute_gauge_values(q, aux)."""
    import pyclaw
    solver, state = _setup(pyclaw)
    dt = solver.dt_initial
    grid = state.grid
    grid.gauge_path = str(tmp_path / '_gauges') + os.sep
    grid.add_gauges([[0.5, 0.25], [1.0, 1.0]])
    solver.compute_gauge_values = lambda q, aux: [q[0], 10 * q[0] + q[1]]
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format = True, None
    claw.tfinal, claw.nout = 8 * dt, 2
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.run()
    files = sorted(os.listdir(grid.gauge_path))
    assert files == ['gauge0.5_0.25.txt', 'gauge1.0_1.0.txt']
    g = np.loadtxt(os.path.join(grid.gauge_path, files[0]))
    assert g.shape == (9, 3)                                   # initial line + 8 steps
    np.testing.assert_allclose(g[:, 0], dt * np.arange(9), rtol=0, atol=1e-15)
    # the gauge cell is floor(coordinate / d), as the reference computes it
    i, j = int(np.floor(0.5 / grid.d[0])), int(np.floor(0.25 / grid.d[1]))
    for k, row in ((0, 0), (1, 4), (2, 8)):
        q = np.asarray(claw.frames[k].q)
        assert g[row, 1] == q[0, i, j] and g[row, 2] == 10 * q[0, i, j] + q[1, i, j]
