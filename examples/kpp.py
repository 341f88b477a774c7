#!/usr/bin/env python
"""The reference's KPP application (apps/kpp/kpp.py) on a USER-SUPPLIED Riemann solver.

The reference links `rpn2_kpp.f rpt2_dummy.f` into classic2.so through RP_SOURCE in the
application's Makefile (apps/kpp/Makefile).  Here the solver is a header
(examples/user_rp/rp_kpp.cuh) that `pyclaw.riemann.from_header` compiles into a variant of
libclawb200.so and binds to `solver.rp`; everything else is the reference script.

    python examples/kpp.py [classic|sharpclaw]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HEADER = os.path.join(ROOT, "examples", "user_rp", "rp_kpp.cuh")


def kpp_solver_descriptor(pyclaw):
    return pyclaw.riemann.from_header(HEADER, name="kpp", meqn=1, mwaves=2, ndims=(2,))


def qinit(state, rad=1.0):
    x = state.grid.x.center
    y = state.grid.y.center
    Y, X = np.meshgrid(y, x)
    r = np.sqrt(X ** 2 + Y ** 2)
    state.q[0, :, :] = 0.25 * np.pi + 3.25 * np.pi * (r <= rad)


def kpp(use_petsc=False, solver_type='classic', mx=200, my=200, tfinal=1.0, nout=10, outdir=None):
    if use_petsc:
        import petclaw as pyclaw
    else:
        import pyclaw
    if solver_type == 'sharpclaw':
        solver = pyclaw.SharpClawSolver2D()
    else:
        solver = pyclaw.ClawSolver2D()
    solver.rp = kpp_solver_descriptor(pyclaw)       # <- the one line the reference has in its Makefile
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    x = pyclaw.Dimension('x', -2.0, 2.0, mx)
    y = pyclaw.Dimension('y', -2.0, 2.0, my)
    state = pyclaw.State(pyclaw.Grid([x, y]), 1)
    qinit(state)
    solver.dim_split = 1
    solver.cfl_max = 1.0
    solver.cfl_desired = 0.9
    solver.mwaves = 2
    solver.limiters = pyclaw.limiters.tvd.minmod
    claw = pyclaw.Controller()
    claw.tfinal = tfinal
    claw.solution = pyclaw.Solution(state)
    claw.solver = solver
    claw.nout = nout
    claw.keep_copy = True
    claw.output_format = None if outdir is None else 'ascii'
    if outdir is not None:
        claw.outdir = outdir
    claw.run()
    return claw


if __name__ == "__main__":
    c = kpp(solver_type=sys.argv[1] if len(sys.argv) > 1 else 'classic')
    q = np.asarray(c.frames[-1].q)
    print("KPP at t = %g: min %.6f max %.6f mean %.6f" % (c.frames[-1].t, q.min(), q.max(), q.mean()))
