"""Regression tests for defects found in review."""
import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu


def _acoustics_solver(pyclaw, lim):
    solver = pyclaw.ClawSolver2D()
    solver.mwaves = 2
    solver.limiters = list(lim)
    solver.dim_split = False
    solver.order_trans = 2
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    return solver


def _state(pyclaw, mx, my, seed):
    x = pyclaw.Dimension('x', -1.0, 1.0, mx)
    y = pyclaw.Dimension('y', -1.0, 1.0, my)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3)
    state.aux_global.update(rho=1.0, bulk=4.0, cc=2.0, zz=2.0)
    state.q[...] = problems.smooth_state("acoustics", (mx, my), seed=seed)
    return state


def _advance(solver, state, nsteps, pyclaw):
    sol = pyclaw.Solution(state)
    solver.setup(sol)
    solver.dt = 0.2 * min(state.grid.d) / 2.0
    for _ in range(nsteps):
        solver.evolve_to_time(sol)
    return np.asarray(sol.state.q).copy()


def test_cuda_graphs_do_not_survive_a_new_setup():
    """One solver object set up on different grids / limiters in turn: every run must equal a
    fresh solver without graph replay (the captured graphs hold the old problem and buffers)."""
    import pyclaw
    reused = _acoustics_solver(pyclaw, [4, 4])
    assert reused.use_cuda_graph
    for k, (mx, my, lim) in enumerate([(64, 48, [4, 4]), (40, 56, [1, 1]), (96, 32, [3, 2]), (64, 48, [4, 4])]):
        reused.limiters = list(lim)
        got = _advance(reused, _state(pyclaw, mx, my, k), 6, pyclaw)
        fresh = _acoustics_solver(pyclaw, lim)
        fresh.use_cuda_graph = False
        want = _advance(fresh, _state(pyclaw, mx, my, k), 6, pyclaw)
        assert np.array_equal(got, want), (mx, my, lim)


def test_cuda_graph_key_covers_boundary_conditions_and_aux():
    """Changing solver.bc_* after the first capture must change what the next steps do."""
    import pyclaw

    def run(graph):
        solver = _acoustics_solver(pyclaw, [4, 4])
        solver.use_cuda_graph = graph
        state = _state(pyclaw, 48, 40, 3)
        sol = pyclaw.Solution(state)
        solver.setup(sol)
        solver.dt = 0.004
        for _ in range(5):
            solver.evolve_to_time(sol)
        solver.bc_lower[0] = pyclaw.BC.reflecting
        solver.bc_upper[1] = pyclaw.BC.periodic
        solver.bc_lower[1] = pyclaw.BC.periodic
        for _ in range(5):
            solver.evolve_to_time(sol)
        return np.asarray(sol.state.q).copy()
    assert np.array_equal(run(True), run(False))


def test_state_reused_with_another_ghost_width():
    """A classic run with fixed dt (no accept hook in round 1) followed by SharpClaw on the same
    State: the ping-pong buffers of the mbc = 2 shape must not come back as spares."""
    import pyclaw
    state = _state(pyclaw, 40, 36, 9)
    s1 = _acoustics_solver(pyclaw, [4, 4])
    s1.dt_variable = False
    sol = pyclaw.Solution(state)
    s1.setup(sol)
    s1.dt = 0.002
    for _ in range(3):
        s1.evolve_to_time(sol)
    q_mid = np.asarray(state.q).copy()
    s2 = pyclaw.SharpClawSolver2D()
    s2.mwaves = 2
    s2.time_integrator = 'SSP33'
    s2.cfl_max, s2.cfl_desired = 0.6, 0.5
    for i in range(2):
        s2.bc_lower[i] = s2.bc_upper[i] = pyclaw.BC.outflow
    s2.setup(sol)
    s2.dt = 0.002
    for _ in range(3):
        s2.evolve_to_time(sol)
    got = np.asarray(state.q).copy()
    fresh = _state(pyclaw, 40, 36, 9)
    fresh.q[...] = q_mid
    s3 = pyclaw.SharpClawSolver2D()
    s3.mwaves = 2
    s3.time_integrator = 'SSP33'
    s3.cfl_max, s3.cfl_desired = 0.6, 0.5
    for i in range(2):
        s3.bc_lower[i] = s3.bc_upper[i] = pyclaw.BC.outflow
    sol3 = pyclaw.Solution(fresh)
    s3.setup(sol3)
    s3.dt = 0.002
    for _ in range(3):
        s3.evolve_to_time(sol3)
    assert np.array_equal(got, np.asarray(fresh.q))


def test_fma_build_is_selectable_and_close():
    """solver.arithmetic = 'fma' runs the -fmad=true build of the same kernels: same step
    sequence, differences at round-off level (the table is in profiles/README.md)."""
    import pyclaw

    def run(arith):
        solver = _acoustics_solver(pyclaw, [4, 4])
        solver.arithmetic = arith
        return _advance(solver, _state(pyclaw, 64, 48, 1), 10, pyclaw)
    a, b = run('strict'), run('fma')
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()
    with pytest.raises(Exception):
        s = _acoustics_solver(pyclaw, [4, 4])
        s.arithmetic = 'fast'
        _advance(s, _state(pyclaw, 16, 16, 1), 1, pyclaw)
