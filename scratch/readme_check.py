import numpy as np, tempfile, os
import pyclaw
solver = pyclaw.ClawSolver2D()
solver.mwaves, solver.limiters = 2, pyclaw.limiters.tvd.MC
solver.dim_split, solver.order_trans = False, 2
for i in range(2):
    solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
grid = pyclaw.Grid([pyclaw.Dimension('x', -1., 1., 256), pyclaw.Dimension('y', -1., 1., 256)])
state = pyclaw.State(grid, 3)
state.aux_global.update(rho=1., bulk=4., cc=2., zz=2.)
Y, X = np.meshgrid(grid.y.center, grid.x.center)
r = np.sqrt(X**2 + Y**2)
state.q[0, :, :] = (np.abs(r - 0.5) <= 0.2) * (1. + np.cos(np.pi * (r - 0.5) / 0.2))
solver.dt_initial = 1e-4
claw = pyclaw.Controller()
claw.solution, claw.solver = pyclaw.Solution(state), solver
claw.outdir = tempfile.mkdtemp()
claw.tfinal, claw.nout, claw.output_format = 0.1, 2, 'ascii'
print(claw.run(), sorted(os.listdir(claw.outdir)))
