"""3-D classic solver: the dimensionally split (step3ds) and the unsplit
(step3 + flux3 with rpt3 / rptt3) kernels through the C ABI against the oracle, and both variants
of the reference's test/acoustics/3d/acoustics.py through ``import pyclaw``: 'hom'
(test_examples.py:474-488, a scalar) and 'het' (test_examples.py:497-514, golden pressure_3D.txt)."""
import os

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


@pytest.mark.parametrize("shape", [(9, 7, 5), (130, 6, 4), (20, 33, 17), (5, 4, 70)])
@pytest.mark.parametrize("order,trans,lim", [(2, 22, [4, 4]), (2, 21, [4, 4]), (2, 20, [2, 3]), (2, 11, [4, 4]),
                                             (1, 11, [0, 0]), (1, 10, [0, 0]), (2, 10, [1, 0]), (2, 0, [4, 4]),
                                             (1, 0, [0, 0])])
def test_step3_unsplit_vs_oracle(shape, order, trans, lim):
    """classic3.step3 for every setting of method(3) that flux3.f:42-68 lists, heterogeneous
    material in all three directions (every aux1 / aux2 / aux3 index choice of rpt3 and rptt3
    matters), bit for bit."""
    mx, my, mz = shape
    mbc = 2
    dx, dy, dz, dt = 0.02, 0.025, 0.03, 0.004
    rng = np.random.RandomState(mx + 10 * my + order + trans)
    pad = (mx + 2 * mbc, my + 2 * mbc, mz + 2 * mbc)
    q = np.asfortranarray(rng.uniform(-1, 1, (4,) + pad))
    aux = np.asfortranarray(np.stack([rng.choice([1.0, 2.0, 3.5], pad), rng.choice([1.0, 2.0], pad)]))
    method = [1, order, trans, 0, 0, 0, 2]
    P = _lib.make_problem(3, 4, 2, mbc, mx, my, dx, dy, po.RP_ACOUSTICS3D_VC, [], method, lim, maux=2)
    qn_o = q.copy("F")
    cfl_o = po.step3(po.RP_ACOUSTICS3D_VC, [], mbc, mx, my, mz, q, qn_o, aux, dx, dy, dz, dt, method, lim)
    qn_g = np.zeros_like(q, order="F")
    cfl_g = ctypes.c_double()
    _lib.call("clawb200_step3_host", ctypes.byref(P), mz, dz, _ptr(q), _ptr(qn_g), _ptr(aux), dt,
              ctypes.byref(cfl_g))
    inner = (slice(None),) + (slice(mbc, -mbc),) * 3   # only the interior is kept (clawpack.py:685)
    assert np.array_equal(qn_g[inner], qn_o[inner]), np.abs(qn_g[inner] - qn_o[inner]).max()
    assert cfl_g.value == cfl_o and cfl_o > 0.05
    assert np.abs(qn_o[inner] - q[inner]).max() > 1e-3


def test_step3_argument_checks():
    P = _lib.make_problem(3, 4, 2, 2, 8, 8, 0.1, 0.1, po.RP_ACOUSTICS3D_VC, [], [1, 1, 22, 0, 0, 0, 2], [4, 4], maux=2)
    q = np.zeros((4, 12, 12, 12), order="F")
    aux = np.ones((2, 12, 12, 12), order="F")
    cfl = ctypes.c_double()
    with pytest.raises(_lib.ClawB200Error, match="method\\[1\\] must be 2"):
        _lib.call("clawb200_step3_host", ctypes.byref(P), 8, 0.1, _ptr(q), _ptr(q.copy("F")), _ptr(aux), 0.01, ctypes.byref(cfl))
    P.method[1], P.method[2] = 2, 12
    with pytest.raises(_lib.ClawB200Error, match="0, 10, 11, 20, 21, 22"):
        _lib.call("clawb200_step3_host", ctypes.byref(P), 8, 0.1, _ptr(q), _ptr(q.copy("F")), _ptr(aux), 0.01, ctypes.byref(cfl))


def _acoustics3d(pyclaw, test, mx, my, mz, tfinal, nout, dim_split=True, order_trans=None, cfl=None):
    solver = pyclaw.ClawSolver3D()
    for i in range(3):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.periodic
        solver.aux_bc_lower[i] = solver.aux_bc_upper[i] = pyclaw.BC.periodic
    solver.dim_split = dim_split
    if order_trans is not None:
        solver.order_trans = order_trans
    if cfl is not None:
        solver.cfl_max, solver.cfl_desired = cfl
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


def test_acoustics3d_heterogeneous_unsplit_golden():
    """test_examples.py:497-514: the 'het' variant (unsplit, order_trans = 22, reflecting lower /
    periodic upper boundaries, 30^3, t = 2) against the reference's golden pressure_3D.txt -- the
    reference asks |diff|_2 < 1e-4; the oracle reproduces every printed digit -- and against the
    oracle driver bit for bit."""
    import pyclaw
    claw, grid, q0, aux = _acoustics3d(pyclaw, 'het', 30, 30, 30, 2.0, 10, dim_split=False)
    assert claw.solver.order_trans == pyclaw.ClawSolver3D.trans_cor
    pfinal = np.asarray(claw.frames[claw.nout].q)[0].reshape(-1)
    gold = np.loadtxt(os.path.join(os.path.dirname(__file__), "golden", "pressure_3D.txt"))
    assert np.linalg.norm(pfinal - gold) < 1e-13 and np.abs(pfinal - gold).max() < 1e-14
    s = po.OracleSolver("classic", 3, po.RP_ACOUSTICS3D_VC, [], 2)
    s.limiters = 4
    s.bc_lower = s.aux_bc_lower = [po.BC_REFLECTING] * 3
    s.bc_upper = s.aux_bc_upper = [po.BC_PERIODIC] * 3
    s.dim_split, s.order_trans = False, 22
    qo = s.run(q0, aux, list(grid.d), 2.0, 10)[-1]
    assert np.array_equal(np.asarray(claw.frames[claw.nout].q), qo)
    assert claw.solver.status['numsteps'] > 0


@pytest.mark.parametrize("order_trans", [0, 11])
def test_acoustics3d_unsplit_other_transverse_settings_vs_oracle(order_trans):
    import pyclaw
    # the donor-cell method (no transverse propagation) is stable up to CFL 0.5 (clawpack.py:588)
    cfl = (0.5, 0.45) if order_trans == 0 else None
    claw, grid, q0, aux = _acoustics3d(pyclaw, 'het', 14, 12, 10, 0.3, 2, dim_split=False,
                                       order_trans=order_trans, cfl=cfl)
    s = po.OracleSolver("classic", 3, po.RP_ACOUSTICS3D_VC, [], 2)
    s.limiters = 4
    s.bc_lower = s.aux_bc_lower = [po.BC_REFLECTING] * 3
    s.bc_upper = s.aux_bc_upper = [po.BC_PERIODIC] * 3
    s.dim_split, s.order_trans = False, order_trans
    if cfl:
        s.cfl_max, s.cfl_desired = cfl
    qo = s.run(q0, aux, list(grid.d), 0.3, 2)[-1]
    assert not np.isnan(qo).any() and np.abs(qo - q0).max() > 0.05
    assert np.array_equal(np.asarray(claw.frames[-1].q), qo)
