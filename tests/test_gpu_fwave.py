"""f-wave path end to end: the reference's two f-wave applications,
apps/elasticity/1d/stegoton/stegoton.py and test/psystem/psystem.py, shortened, run through
``import pyclaw`` with ``solver.fwave = True`` against the oracle driver.

The reference holds no golden data for either (test_examples.py:440-444 returns True), and the
Riemann solvers are external (clawpack/riemann, un-vendored): parity here is GPU vs oracle --
bit for bit with the linear stress law, to rounding error with the exponential law (exp()
comes from the CUDA math library on the GPU and from libm in the oracle)."""
import numpy as np
import pytest

from oracle import pyclaw_oracle as po

pytestmark = pytest.mark.gpu


def _stegoton_setaux(x, rhoB=4, KB=4, rhoA=1, KA=1, alpha=0.5):
    aux = np.empty([3, len(x)], order='F')
    xfrac = x - np.floor(x)
    aux[0, :] = rhoA * (xfrac < alpha) + rhoB * (xfrac >= alpha)
    aux[1, :] = KA * (xfrac < alpha) + KB * (xfrac >= alpha)
    aux[2, :] = 0.
    return aux


def _stegoton_q(x, aux, law, xupper):
    q = np.zeros((2, len(x)), order='F')
    sigma = 1.0 * np.exp(-((x - xupper / 2.) / 5.) ** 2.)
    q[0] = np.log(sigma + 1.) / aux[1] if law == 2 else sigma / aux[1]
    return q


@pytest.mark.parametrize("law", [1, 2])
@pytest.mark.parametrize("solver_type", ['classic', 'sharpclaw'])
def test_stegoton_fwave_vs_oracle(law, solver_type):
    import pyclaw
    xupper, cellsperlayer = 60.0, 6
    mx = int(round(xupper)) * cellsperlayer
    if solver_type == 'classic':
        solver = pyclaw.ClawSolver1D()
    else:                                   # stegoton.py:85-86,137-139
        solver = pyclaw.SharpClawSolver1D()
        solver.lim_type, solver.char_decomp = 2, 0
    solver.kernel_language = 'Fortran'
    solver.bc_lower[0] = solver.bc_upper[0] = pyclaw.BC.periodic
    solver.aux_bc_lower[0] = solver.aux_bc_upper[0] = pyclaw.BC.periodic
    solver.fwave = True
    solver.mwaves = 2
    grid = pyclaw.Grid(pyclaw.Dimension('x', 0.0, xupper, mx))
    state = pyclaw.State(grid, 2)
    state.aux_global.update(KA=1.0, KB=4.0, rhoA=1.0, rhoB=4.0, stress_law=law)
    xc = grid.x.center
    aux = _stegoton_setaux(xc)
    state.aux = aux
    q0 = _stegoton_q(xc, aux, law, xupper)
    state.q[...] = q0
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format, claw.tfinal, claw.nout = True, None, 8.0, 2
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    status = claw.run()
    qg = np.asarray(claw.frames[-1].q)

    s = po.OracleSolver(solver_type, 1, po.RP_NEL_FWAVE, [float(law)], 2)
    s.bc_lower = s.bc_upper = [po.BC_PERIODIC]
    s.aux_bc_lower = s.aux_bc_upper = [po.BC_PERIODIC]
    qo = s.run(q0, aux, [grid.d[0]], 8.0, 2)[-1]
    assert not np.isnan(qo).any() and np.abs(qo - q0).max() > 0.05
    if law == 1:
        assert np.array_equal(qg, qo)
    else:
        assert np.allclose(qg, qo, rtol=0, atol=1e-12)
    # strain and momentum are conserved by the f-wave scheme on the periodic domain
    assert abs(qg[0].sum() - q0[0].sum()) < 1e-11 and abs(qg[1].sum()) < 1e-11


def _psystem_setaux(x, y, lin):
    E1 = p1 = 1.
    E2 = p2 = 4.
    xfrac = x - np.floor(x)
    yfrac = y - np.floor(y)
    yy, xx = np.meshgrid(yfrac, xfrac)
    a = (xx <= 0.5) * (yy <= 0.5) + (xx > 0.5) * (yy > 0.5)
    b = (xx > 0.5) * (yy <= 0.5) + (xx <= 0.5) * (yy > 0.5)
    aux = np.empty((4, len(x), len(y)), order='F')
    aux[0] = p1 * a + p2 * b
    aux[1] = E1 * a + E2 * b
    aux[2] = lin
    return aux


@pytest.mark.parametrize("lin,dim_split", [(1, False), (2, False), (1, True)])
def test_psystem_fwave_vs_oracle(lin, dim_split):
    import pyclaw
    Ng = 8
    mx = my = 3 * Ng
    solver = pyclaw.ClawSolver2D()
    solver.mwaves = 2
    solver.limiters = pyclaw.limiters.tvd.superbee
    solver.bc_lower[0] = solver.bc_lower[1] = pyclaw.BC.reflecting
    solver.bc_upper[0] = solver.bc_upper[1] = pyclaw.BC.outflow
    solver.aux_bc_lower[0] = solver.aux_bc_lower[1] = pyclaw.BC.reflecting
    solver.aux_bc_upper[0] = solver.aux_bc_upper[1] = pyclaw.BC.outflow
    solver.fwave = True
    solver.cfl_max, solver.cfl_desired = 1.0, 0.9
    solver.dim_split = dim_split
    solver.order_trans = 2
    grid = pyclaw.Grid([pyclaw.Dimension('x', 0.25, 3.25, mx), pyclaw.Dimension('y', 0.25, 3.25, my)])
    state = pyclaw.State(grid, 3, 4)
    aux = _psystem_setaux(grid.x.center, grid.y.center, lin)
    yy, xx = np.meshgrid(grid.y.center, grid.x.center)
    s0 = 5. * np.exp(-(xx - 0.25) ** 2 / (2 * 0.5) - (yy - 0.25) ** 2 / (2 * 0.5))
    q0 = np.zeros((3, mx, my), order='F')
    q0[0] = s0 / aux[1] if lin == 1 else np.log(s0 + 1) / aux[1]
    aux[3] = q0[0]
    state.aux = aux
    state.q[...] = q0
    solver.dt_initial = 0.01
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format, claw.tfinal, claw.nout = True, None, 0.5, 2
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.run()
    qg = np.asarray(claw.frames[-1].q)

    s = po.OracleSolver("classic", 2, po.RP_PSYSTEM, [], 2)
    s.limiters = 2
    s.bc_lower = [po.BC_REFLECTING] * 2
    s.bc_upper = [po.BC_OUTFLOW] * 2
    s.aux_bc_lower = [po.BC_REFLECTING] * 2
    s.aux_bc_upper = [po.BC_OUTFLOW] * 2
    s.dim_split, s.order_trans = dim_split, 2
    s.dt_initial = 0.01
    qo = s.run(q0, aux, list(grid.d), 0.5, 2)[-1]
    assert not np.isnan(qo).any() and np.abs(qo - q0).max() > 0.05
    if lin == 1:
        assert np.array_equal(qg, qo)
    else:
        assert np.allclose(qg, qo, rtol=0, atol=1e-11)


def test_fwave_flag_must_match_the_solver():
    import pyclaw
    solver = pyclaw.ClawSolver1D()
    solver.bc_lower[0] = solver.bc_upper[0] = pyclaw.BC.periodic
    solver.mwaves = 2
    solver.rp = pyclaw.riemann.acoustics
    solver.fwave = True
    state = pyclaw.State(pyclaw.Grid(pyclaw.Dimension('x', 0., 1., 50)), 2)
    state.aux_global.update(rho=1., bulk=1., cc=1., zz=1.)
    with pytest.raises(Exception, match="f-?waves|waves"):
        solver.setup(pyclaw.Solution(state))
