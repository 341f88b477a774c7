"""
Parity AT THE BASELINE SIZES (BASELINE.json configs 1-4): one full time step of each
configuration on the grid the benchmark times -- 4096^2 acoustics, 4096^2 and 8192^2 Euler,
4096^2 SharpClaw shallow water (SSP33), 4096 x 2048 shallow water on the sphere -- through
  * the f2py-shaped host entry points of the C ABI (`clawb200_step2_host`, the pipelined slab
    path behind bench.py's `e2e` figure; `clawb200_sharpclaw_dq_host`), and
  * the device-resident path of the PyClaw API (what `value` times),
against the CPU oracle run on all host cores (y-slab threads; identical results to the serial
oracle, tests/test_oracle_golden.py).  float64, BIT FOR BIT (np.array_equal), CFL number included.
"""
import ctypes
import os

import numpy as np
import pytest

import problems
from oracle import pyclaw_oracle as po
from pyclaw_b200 import _lib

pytestmark = pytest.mark.gpu
NTHREADS = max(2, os.cpu_count() or 2)

RPS = {
    "acoustics": (1, [1.0, 4.0, 2.0, 2.0], 3, 2, [4, 4]),
    "euler": (3, [1.4, 0.4], 5, 5, [4, 4, 4, 4, 2]),
    "shallow": (4, [1.0], 3, 3, [4, 4, 4]),
}


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _inner(mbc):
    return (slice(None), slice(mbc, -mbc), slice(mbc, -mbc))


def _same(a, b, what):
    # chunked comparison: no 2.7-GB temporaries
    for m in range(a.shape[0]):
        if not np.array_equal(a[m], b[m]):
            d = np.abs(a[m] - b[m])
            raise AssertionError("%s: component %d differs, max |diff| = %g at %s"
                                 % (what, m, d.max(), np.unravel_index(d.argmax(), d.shape)))


def _classic_api_step(rp, q_interior, dx, dy, dt, lim, aux_global):
    """One unsplit step through the PyClaw API with q resident on the device (outflow ghost
    cells filled by bc_kernel)."""
    import pyclaw
    mx, my = q_interior.shape[1:]
    x = pyclaw.Dimension('x', 0.0, dx * mx, mx)
    y = pyclaw.Dimension('y', 0.0, dy * my, my)
    state = pyclaw.State(pyclaw.Grid([x, y]), q_interior.shape[0])
    state.aux_global.update(aux_global)
    state.q[...] = q_interior
    solver = pyclaw.ClawSolver2D()
    solver.mwaves = len(lim)
    solver.limiters = list(lim)
    solver.dim_split = False
    solver.order_trans = 2
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    sol = pyclaw.Solution(state)
    solver.setup(sol)
    # the reference computes dx as (upper - lower) / n; use the very same number in the oracle
    solver.dt = dt
    solver.step(sol)
    return np.asarray(sol.state.q), solver.cfl.get_cached_max(), state.grid.d


@pytest.mark.parametrize("rp,n", [("acoustics", 4096), ("euler", 4096), ("euler", 8192)])
def test_unsplit_step_at_baseline_size(rp, n):
    rp_id, params, meqn, mwaves, lim = RPS[rp]
    mbc = 2
    mx = my = n
    dx = dy = 2.0 / n
    dt = 0.12 * dx
    method = [1, 2, 2, 0, 0, 0, 0]
    q = np.empty((meqn, mx + 2 * mbc, my + 2 * mbc), order="F")
    q[_inner(mbc)] = problems.big_state(rp, (mx, my), seed=n % 97)
    po.fill_bcs(q, mbc, [po.BC_OUTFLOW] * 2, [po.BC_OUTFLOW] * 2)
    # ---- oracle on all host cores ----
    qn_o = q.copy("F")
    cfl_o = po.step2_slabs(rp_id, params, mbc, mx, my, q, qn_o, None, dx, dy, dt, method, lim, NTHREADS, False)
    assert 0.05 < cfl_o < 1.0 and np.isfinite(qn_o[0]).all()
    # ---- host-buffer C ABI (slab pipeline: H2D, layout, sweeps, layout, D2H) ----
    qn_g = q.copy("F")
    cfl_g = ctypes.c_double()
    P = _lib.make_problem(2, meqn, mwaves, mbc, mx, my, dx, dy, rp_id, params, method, lim)
    _lib.call("clawb200_step2_host", ctypes.byref(P), _ptr(q), _ptr(qn_g), None, dt, ctypes.byref(cfl_g))
    _same(qn_g[_inner(mbc)], qn_o[_inner(mbc)], "clawb200_step2_host %s %d^2" % (rp, n))
    assert cfl_g.value == cfl_o
    del qn_g
    # ---- device-resident path through the API ----
    names = {"acoustics": dict(rho=1.0, bulk=4.0, cc=2.0, zz=2.0), "euler": dict(gamma=1.4, gamma1=0.4)}[rp]
    got, cfl_a, d = _classic_api_step(rp, q[_inner(mbc)], dx, dy, dt, lim, names)
    assert d[0] == dx and d[1] == dy
    _same(got, qn_o[_inner(mbc)], "ClawSolver2D.step %s %d^2" % (rp, n))
    assert cfl_a == cfl_o


def test_sharpclaw_ssp33_step_at_baseline_size():
    """BASELINE config 3 at 4096^2: the dq kernel through the host entry point, and one SSP33
    step (three fused stage kernels) through the API."""
    import pyclaw
    n, mbc = 4096, 3
    rp_id, params, meqn, mwaves, _ = RPS["shallow"]
    mx = my = n
    dx = dy = 5.0 / n
    dt = 0.1 * dx
    qi = problems.big_state("shallow", (mx, my), seed=11)
    bl, bu = [po.BC_OUTFLOW] * 2, [po.BC_REFLECTING] * 2
    q = np.empty((meqn, mx + 2 * mbc, my + 2 * mbc), order="F")
    q[_inner(mbc)] = qi
    po.fill_bcs(q, mbc, bl, bu)
    dq_o, cfl_o = po.sc_flux2(rp_id, params, mwaves, mbc, mx, my, q, dx, dy, dt, po.WENO_PYWENO_F32, NTHREADS)
    dq_g = np.zeros_like(q, order="F")
    cfl_g = ctypes.c_double()
    P = _lib.make_problem(2, meqn, mwaves, mbc, mx, my, dx, dy, rp_id, params, weno_variant=0)
    _lib.call("clawb200_sharpclaw_dq_host", ctypes.byref(P), _ptr(q), _ptr(dq_g), None, dt, ctypes.byref(cfl_g))
    _same(dq_g[_inner(mbc)], dq_o[_inner(mbc)], "clawb200_sharpclaw_dq_host 4096^2")
    assert cfl_g.value == cfl_o
    del dq_g, dq_o, q
    # one SSP33 step, oracle vs API
    s = po.OracleSolver("sharpclaw", 2, rp_id, params, 3)
    s.time_integrator = "SSP33"
    s.bc_lower, s.bc_upper = bl, bu
    s.nthreads = NTHREADS
    s.setup(qi, None, [dx, dy])
    s.dt = dt
    st = {"q": qi.copy("F"), "t": 0.0}
    s.step(st)
    x = pyclaw.Dimension('x', 0.0, dx * mx, mx)
    y = pyclaw.Dimension('y', 0.0, dy * my, my)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3)
    assert state.grid.d[0] == dx
    state.aux_global['grav'] = 1.0
    state.q[...] = qi
    solver = pyclaw.SharpClawSolver2D()
    solver.mwaves = 3
    solver.time_integrator = 'SSP33'
    solver.bc_lower[0] = solver.bc_lower[1] = pyclaw.BC.outflow
    solver.bc_upper[0] = solver.bc_upper[1] = pyclaw.BC.reflecting
    sol = pyclaw.Solution(state)
    solver.setup(sol)
    solver.dt = dt
    solver.step(sol)
    _same(np.asarray(sol.state.q), st["q"], "SharpClawSolver2D SSP33 step 4096^2")
    assert solver.cfl.get_cached_max() == s.cfl


def test_sphere_hyperbolic_step_at_baseline_size():
    """BASELINE config 4 at 4096 x 2048: step2qcor sweeps with 16 aux components, capacity
    function, periodic x and pole-fold y ghost cells.  The application's source term is a
    pointwise kernel checked separately (test_sphere_src2_kernel_vs_tensor_ops_vs_oracle); the
    same aux / q arrays are handed to both sides, so the comparison is bit for bit."""
    import pyclaw
    from pyclaw_b200.apps import shallow_sphere as app
    mx, my, mbc = 4096, 2048, 2
    state, solver = app.setup(pyclaw, mx, my)
    solver.step_src = None
    sol = pyclaw.Solution(state)
    solver.setup(sol)
    dx, dy = state.grid.d
    q0 = np.asfortranarray(np.asarray(state.q))
    auxbc = np.asfortranarray(np.asarray(solver.auxbc))
    dt = 0.02 * dx
    solver.dt = dt
    solver.step(sol)
    got = np.asarray(sol.state.q)
    cfl_g = solver.cfl.get_cached_max()

    qbc = np.zeros((4, mx + 2 * mbc, my + 2 * mbc), order="F")
    qbc[_inner(mbc)] = q0
    po.fill_bcs(qbc, mbc, [po.BC_PERIODIC, po.BC_CUSTOM], [po.BC_PERIODIC, po.BC_CUSTOM],
                problems.sphere_qbc_lower_y, problems.sphere_qbc_upper_y)
    qn = qbc.copy("F")
    method = [1, 2, 2, 0, 0, 1, 16]
    cfl_o = po.step2_slabs(po.RP_SPHERE, [problems.SPHERE_G], mbc, mx, my, qbc, qn, auxbc, dx, dy, dt,
                           method, [4, 4, 4], NTHREADS, False)
    assert 0.01 < cfl_o < 1.0
    _same(got, qn[_inner(mbc)], "sphere step2qcor 4096x2048")
    assert cfl_g == cfl_o
