"""3-D classic solver, dimensional splitting (the step3ds half): kernels
through the C ABI against the oracle, and the reference's test/acoustics/3d/acoustics.py
('hom' variant, test_examples.py:474-488) through ``import pyclaw``."""
import ctypes

import numpy as np
import pytest

from oracle import pyclaw_oracle as po
from pyclaw_b200 import _lib

pytestmark = pytest.mark.gpu


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("shape", [(9, 7, 5), (130, 6, 4), (20, 33, 17), (5, 4, 70)])
@pytest.mark.parametrize("order,lim", [(2, [4, 4]), (1, [0, 0]), (2, [2, 3])])
def test_step3ds_vs_oracle(shape, order, lim):
    mx, my, mz = shape
    mbc = 2
    dx, dy, dz, dt = 0.02, 0.025, 0.03, 0.004
    rng = np.random.RandomState(mx + 10 * my + order)
    pad = (mx + 2 * mbc, my + 2 * mbc, mz + 2 * mbc)
    q = np.asfortranarray(rng.uniform(-1, 1, (4,) + pad))
    aux = np.asfortranarray(np.stack([rng.choice([1.0, 2.0, 3.5], pad), rng.choice([1.0, 2.0], pad)]))
    method = [1, order, -1, 0, 0, 0, 2]
    P = _lib.make_problem(3, 4, 2, mbc, mx, my, dx, dy, po.RP_ACOUSTICS3D_VC, [], method, lim, maux=2)
    cfl_g = ctypes.c_double()
    for idir in (1, 2, 3):
        qn_o = q.copy("F")
        cfl_o = po.step3ds(po.RP_ACOUSTICS3D_VC, [], mbc, mx, my, mz, q, qn_o, aux, dx, dy, dz, dt, method, lim, idir)
        qn_g = np.zeros_like(q, order="F")
        _lib.call("clawb200_step3ds_host", ctypes.byref(P), mz, dz, _ptr(q), _ptr(qn_g), _ptr(aux), dt, idir,
                  ctypes.byref(cfl_g))
        # every cell: swept cells updated, all others equal to qold
        assert np.array_equal(qn_g, qn_o), (idir, np.abs(qn_g - qn_o).max())
        assert cfl_g.value == cfl_o and cfl_o > 0.05
        assert np.abs(qn_o - q).max() > 1e-3


def _acoustics3d(pyclaw, test, mx, my, mz, tfinal, nout):
    solver = pyclaw.ClawSolver3D()
    for i in range(3):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.periodic
        solver.aux_bc_lower[i] = solver.aux_bc_upper[i] = pyclaw.BC.periodic
    solver.dim_split = True
    if test == 'hom':
        zr = cr = 1.0
    else:
        for i in range(3):
            solver.bc_lower[i] = solver.aux_bc_lower[i] = pyclaw.BC.reflecting
        zr = cr = 2.0
    solver.mwaves = 2
    solver.limiters = pyclaw.limiters.tvd.MC
    grid = pyclaw.Grid([pyclaw.Dimension('x', -1.0, 1.0, mx), pyclaw.Dimension('y', -1.0, 1.0, my),
                        pyclaw.Dimension('z', -1.0, 1.0, mz)])
    state = pyclaw.State(grid, 4, 2)
    grid.compute_c_center()
    X, Y, Z = grid._c_center
    aux = np.empty((2, mx, my, mz), order='F')
    aux[0] = 1.0 * (X < 0.) + zr * (X >= 0.)
    aux[1] = 1.0 * (X < 0.) + cr * (X >= 0.)
    state.aux = aux
    q0 = np.zeros((4, mx, my, mz), order='F')
    x0 = -0.5
    if test == 'hom':
        r = np.sqrt((X - x0) ** 2)
        q0[0] = (np.abs(r) <= 0.2) * (1. + np.cos(np.pi * r / 0.2))
    else:
        r = np.sqrt((X - x0) ** 2 + Y ** 2 + Z ** 2)
        q0[0] = (np.abs(r - 0.3) <= 0.1) * (1. + np.cos(np.pi * (r - 0.3) / 0.1))
    state.q[...] = q0
    claw = pyclaw.Controller()
    claw.keep_copy, claw.output_format = True, None
    claw.solution, claw.solver = pyclaw.Solution(state), solver
    claw.tfinal, claw.nout = tfinal, nout
    claw.run()
    return claw, grid, q0, aux


def test_acoustics3d_homogeneous_reference_scalar():
    """test_examples.py:481-488: |final - initial|_1 dx dy dz = 0.00286 +- 1e-4 after one period."""
    import pyclaw
    claw, grid, q0, aux = _acoustics3d(pyclaw, 'hom', 256, 4, 4, 2.0, 10)
    pinitial = np.asarray(claw.frames[0].q)[0].reshape(-1)
    pfinal = np.asarray(claw.frames[claw.nout].q)[0].reshape(-1)
    final_difference = np.prod(grid.d) * np.linalg.norm(pfinal - pinitial, ord=1)
    assert abs(final_difference - 0.00286) < 1e-4, final_difference
    # and the oracle driver gives the same field bit for bit
    s = po.OracleSolver("classic", 3, po.RP_ACOUSTICS3D_VC, [], 2)
    s.limiters = 4
    s.bc_lower = s.bc_upper = [po.BC_PERIODIC] * 3
    s.aux_bc_lower = s.aux_bc_upper = [po.BC_PERIODIC] * 3
    s.dim_split = True
    qo = s.run(q0, aux, list(grid.d), 2.0, 10)[-1]
    assert np.array_equal(np.asarray(claw.frames[claw.nout].q), qo)


def test_acoustics3d_heterogeneous_dimsplit_vs_oracle():
    import pyclaw
    claw, grid, q0, aux = _acoustics3d(pyclaw, 'het', 20, 18, 16, 0.4, 2)
    s = po.OracleSolver("classic", 3, po.RP_ACOUSTICS3D_VC, [], 2)
    s.limiters = 4
    s.bc_lower = s.aux_bc_lower = [po.BC_REFLECTING] * 3
    s.bc_upper = s.aux_bc_upper = [po.BC_PERIODIC] * 3
    s.dim_split = True
    qo = s.run(q0, aux, list(grid.d), 0.4, 2)[-1]
    qg = np.asarray(claw.frames[-1].q)
    assert not np.isnan(qo).any() and np.abs(qo - q0).max() > 0.05
    assert np.array_equal(qg, qo)


def test_unsplit_3d_is_refused():
    import pyclaw
    solver = pyclaw.ClawSolver3D()
    solver.dim_split = False
    solver.mwaves = 2
    grid = pyclaw.Grid([pyclaw.Dimension('x', 0., 1., 8), pyclaw.Dimension('y', 0., 1., 8),
                        pyclaw.Dimension('z', 0., 1., 8)])
    state = pyclaw.State(grid, 4, 2)
    with pytest.raises(NotImplementedError):
        solver.setup(pyclaw.Solution(state))
