"""The Riemann-solver plugin seam on the GPU: the KPP solver of examples/user_rp/rp_kpp.cuh
(the reference's apps/kpp links rpn2_kpp.f + rpt2_dummy.f through RP_SOURCE)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def _hits(a, b, p):
    return np.floor((b - p) / (2 * np.pi)) >= np.ceil((a - p) / (2 * np.pi))


def _kpp_numpy(ixy, ul, ur):
    """The formulas of rp_kpp.cuh in numpy."""
    a, b = np.minimum(ul, ur), np.maximum(ul, ur)
    if ixy == 1:
        fl, fr = np.sin(ul), np.sin(ur)
        smax = np.where(_hits(a, b, 0.0), 1.0, np.maximum(np.cos(a), np.cos(b)))
        smin = np.where(_hits(a, b, np.pi), -1.0, np.minimum(np.cos(a), np.cos(b)))
    else:
        fl, fr = np.cos(ul), np.cos(ur)
        smax = np.where(_hits(a, b, 1.5 * np.pi), 1.0, np.maximum(-np.sin(a), -np.sin(b)))
        smin = np.where(_hits(a, b, 0.5 * np.pi), -1.0, np.minimum(-np.sin(a), -np.sin(b)))
    with np.errstate(invalid="ignore", divide="ignore"):
        um = np.where(smax > smin, (smax * ur - smin * ul - (fr - fl)) / (smax - smin), ul)
    w1, w2 = um - ul, ur - um
    amdq = np.minimum(smin, 0) * w1 + np.minimum(smax, 0) * w2
    apdq = np.maximum(smin, 0) * w1 + np.maximum(smax, 0) * w2
    return np.stack([w1, w2]), np.stack([smin, smax]), amdq, apdq, fl, fr


@pytest.mark.parametrize("ixy", [1, 2])
def test_user_solver_pointwise_contract(ixy):
    """solver.rp as a callable, the reference's Python contract (doc/rp.rst): waves, speeds and
    fluctuations of the compiled-in user solver against the same formulas in numpy; the waves sum
    to the jump and the fluctuations to the flux difference (HLL is conservative)."""
    import torch
    import pyclaw
    import kpp
    rs = kpp.kpp_solver_descriptor(pyclaw)
    rng = np.random.RandomState(ixy)
    n = 5000
    ul, ur = rng.uniform(0, 4 * np.pi, n), rng.uniform(0, 4 * np.pi, n)
    ur[:200] = ul[:200]
    ql = torch.as_tensor(ul[None, :], device="cuda")
    qr = torch.as_tensor(ur[None, :], device="cuda")
    wave, s, amdq, apdq = [t.cpu().numpy() for t in rs(ql, qr, None, None, {}, ixy=ixy)]
    w_np, s_np, am_np, ap_np, fl, fr = _kpp_numpy(ixy, ul, ur)
    assert np.abs(wave[0] - w_np).max() < 1e-12 and np.abs(s - s_np).max() < 1e-14
    assert np.abs(amdq[0] - am_np).max() < 1e-12 and np.abs(apdq[0] - ap_np).max() < 1e-12
    assert np.abs(wave[0].sum(axis=0) - (ur - ul)).max() < 1e-13
    assert np.abs(amdq[0] + apdq[0] - (fr - fl)).max() < 1e-12


@pytest.mark.parametrize("solver_type", ["classic", "sharpclaw"])
def test_kpp_application_runs_on_the_user_solver(solver_type):
    import kpp
    c = kpp.kpp(solver_type=solver_type, mx=80, my=80, tfinal=0.4, nout=2)
    q0 = np.asarray(c.frames[0].q)
    q = np.asarray(c.frames[-1].q)
    assert np.isfinite(q).all() and c.frames[-1].t == 0.4
    # scalar conservation law, outflow boundaries not reached by t = 0.4: mass is conserved
    assert abs(q.sum() - q0.sum()) < 1e-9 * abs(q0.sum())
    # the solution stays (essentially) inside the initial range [pi/4, 3.5 pi]
    assert q.min() > 0.25 * np.pi - 0.05 and q.max() < 3.5 * np.pi + 0.05
    # and it did something: the rotating-wave structure has developed
    assert np.abs(q - q0).max() > 1.0


def test_user_solver_equals_builtin_when_it_is_the_same_solver():
    """A user header that restates a built-in solver (2-D advection) gives bit-identical runs."""
    import pyclaw
    hdr = os.path.join(ROOT, "examples", "user_rp", "rp_advection_user.cuh")
    rs = pyclaw.riemann.from_header(hdr, name="advuser", meqn=1, mwaves=1, ndims=(2,), param_names=["u", "v"])

    def run(rp):
        x = pyclaw.Dimension('x', 0.0, 1.0, 64)
        y = pyclaw.Dimension('y', 0.0, 1.0, 48)
        state = pyclaw.State(pyclaw.Grid([x, y]), 1)
        state.aux_global.update(u=0.7, v=-0.4)
        X, Y = np.meshgrid(state.grid.x.center, state.grid.y.center, indexing="ij")
        state.q[0] = np.exp(-60 * ((X - 0.4) ** 2 + (Y - 0.6) ** 2))
        solver = pyclaw.ClawSolver2D()
        solver.rp = rp
        solver.mwaves = 1
        solver.limiters = [4]
        solver.dim_split = False
        solver.order_trans = 2
        for i in range(2):
            solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.periodic
        claw = pyclaw.Controller()
        claw.tfinal, claw.nout, claw.keep_copy, claw.output_format = 0.3, 1, True, None
        claw.solution, claw.solver = pyclaw.Solution(state), solver
        claw.run()
        return np.asarray(claw.frames[-1].q)
    assert np.array_equal(run(rs), run(pyclaw.riemann.advection))
