import sys, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import problems
from oracle import pyclaw_oracle as po
import pyclaw

pb = problems.shallow2d(60, 60)
ti = 'SSP33'
s = po.OracleSolver('sharpclaw', 2, po.RP_SHALLOW, [1.0], 3)
s.bc_lower = [po.BC_OUTFLOW, po.BC_OUTFLOW]; s.bc_upper = [po.BC_REFLECTING, po.BC_REFLECTING]
s.time_integrator = ti
hist_o = []
orig = s.step
def step_o(state):
    r = orig(state); hist_o.append((state['t'], s.dt, s.cfl, state['q'].copy())); return r
s.step = step_o
fo = s.run(pb['q'], None, pb['d'], 1.0, 2)

solver = pyclaw.SharpClawSolver2D()
solver.mwaves = 3; solver.time_integrator = ti
solver.bc_lower[0] = pyclaw.BC.outflow; solver.bc_upper[0] = pyclaw.BC.reflecting
solver.bc_lower[1] = pyclaw.BC.outflow; solver.bc_upper[1] = pyclaw.BC.reflecting
x = pyclaw.Dimension('x', -2.5, 2.5, 60); y = pyclaw.Dimension('y', -2.5, 2.5, 60)
state = pyclaw.State(pyclaw.Grid([x, y]), 3); state.aux_global['grav'] = 1.0
state.q[...] = pb['q']
hist_g = []
orig_g = solver.step
def step_g(sol):
    r = orig_g(sol); hist_g.append((sol.t, solver.dt, solver.cfl.get_cached_max(), np.asarray(sol.state.q).copy())); return r
solver.step = step_g
claw = pyclaw.Controller(); claw.tfinal = 1.0; claw.keep_copy = True; claw.output_format = None
claw.solution = pyclaw.Solution(state); claw.solver = solver; claw.nout = 2
claw.run()
print(len(hist_o), len(hist_g))
for i, (a, b) in enumerate(zip(hist_o, hist_g)):
    d = np.abs(a[3]-b[3])
    idx = np.argwhere(d > 0)
    print(i, a[:3], b[:3], 'qdiff', d.max(), len(idx), idx[:6].tolist())
