import sys, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import problems
from oracle import pyclaw_oracle as po
import pyclaw

pb = problems.shallow2d(60, 60)
for ti in ['Euler', 'SSP33', 'SSP104']:
    s = po.OracleSolver('sharpclaw', 2, po.RP_SHALLOW, [1.0], 3)
    s.bc_lower = [po.BC_OUTFLOW, po.BC_OUTFLOW]; s.bc_upper = [po.BC_REFLECTING, po.BC_REFLECTING]
    s.time_integrator = ti
    q0 = pb['q'].copy('F')
    s.setup(q0, None, pb['d'])
    s.dt = 0.01
    st = {'q': q0.copy('F'), 't': 0.0}
    for k in range(3):
        s.step(st)
    solver = pyclaw.SharpClawSolver2D()
    solver.mwaves = 3; solver.time_integrator = ti
    solver.bc_lower[0] = pyclaw.BC.outflow; solver.bc_upper[0] = pyclaw.BC.reflecting
    solver.bc_lower[1] = pyclaw.BC.outflow; solver.bc_upper[1] = pyclaw.BC.reflecting
    x = pyclaw.Dimension('x', -2.5, 2.5, 60); y = pyclaw.Dimension('y', -2.5, 2.5, 60)
    state = pyclaw.State(pyclaw.Grid([x, y]), 3); state.aux_global['grav'] = 1.0
    state.q[...] = pb['q']
    sol = pyclaw.Solution(state)
    solver.setup(sol); solver.dt = 0.01
    for k in range(3):
        solver.step(sol)
        print(ti, k, 'cfl', solver.cfl.get_cached_max())
    g = np.asarray(sol.state.q)
    print(ti, 'oracle cfl', s.cfl, 'maxdiff', np.abs(g - st['q']).max(), 'nz', (g != st['q']).sum())
