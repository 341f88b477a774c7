"""Times one SharpClaw SSP33 step of the Euler shock-bubble problem (not a bench workload)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pyclaw, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for kind in ("shallow", "euler"):
    solver = pyclaw.SharpClawSolver2D()
    solver.time_integrator = 'SSP33'
    solver.cfl_max, solver.cfl_desired = 0.6, 0.5
    for i in range(2):
        solver.bc_lower[i] = solver.bc_upper[i] = pyclaw.BC.outflow
    grid = pyclaw.Grid([pyclaw.Dimension('x', 0., 2., n), pyclaw.Dimension('y', 0., 2., n)])
    if kind == "euler":
        pb = problems.shockbubble(n, n, xupper=2.0, yupper=2.0)
        state = pyclaw.State(grid, 5); solver.mwaves = 5
        state.aux_global.update(gamma=1.4, gamma1=0.4)
    else:
        pb = problems.shallow2d(n, n)
        state = pyclaw.State(grid, 3); solver.mwaves = 3
        state.aux_global['grav'] = 1.0
    state.q[...] = pb["q"]
    solver.dt_initial = 0.1 * grid.d[0]
    sol = pyclaw.Solution(state)
    solver.setup(sol); solver.dt = solver.dt_initial
    for _ in range(3): solver.evolve_to_time(sol)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): solver.evolve_to_time(sol)
    torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 5
    print("%s %d^2: %.2f ms/step, %.3f G cell-updates/s" % (kind, n, el * 1e3, n * n / el / 1e9), flush=True)
