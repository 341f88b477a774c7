"""Host-side logic of the PyClaw-compatible API, on CPU tensors (no kernels): data model,
option handling and the dt / accept-reject loop of Solver.evolve_to_time
(src/pyclaw/solver.py:602-717)."""
import copy

import numpy as np
import pytest
import torch

import pyclaw
from pyclaw_b200 import riemann
from pyclaw_b200.solver import Solver


def _state2d(mx=6, my=4, meqn=3, maux=0):
    x = pyclaw.Dimension('x', 0., 3., mx)
    y = pyclaw.Dimension('y', -1., 1., my)
    return pyclaw.State(pyclaw.Grid([x, y]), meqn, maux, device='cpu')


def test_dimension_and_grid():
    x = pyclaw.Dimension('x', 0., 3., 6)
    assert x.d == 0.5 and x.ng == 6 and x.nstart == 0 and x.nend == 6
    assert np.allclose(x.center, 0.25 + 0.5 * np.arange(6)) and len(x.edge) == 7
    g = pyclaw.Grid([x, pyclaw.Dimension('y', -1., 1., 4)])
    assert g.ndim == 2 and g.n == [6, 4] and g.d == [0.5, 0.5] and g.name == ['x', 'y']
    X, Y = g.c_center
    assert X.shape == (6, 4) and X[2, 0] == x.center[2] and Y[0, 3] == g.y.center[3]
    assert pyclaw.Dimension(0., 1., 10).name == 'x'


def test_state_views_and_layout():
    s = _state2d()
    assert tuple(s.q.shape) == (3, 6, 4) and s.meqn == 3 and s.maux == 0 and s.aux is None
    s.q[0, :, :] = np.arange(24.).reshape(6, 4)          # numpy assignment, reference idiom
    s.q[1, :, :] = 0.
    s.set_mbc(2)
    assert tuple(s.q.shape) == (3, 6, 4)
    assert float(s.q[0, 5, 3]) == 23.0
    # storage is [m][j][i] with i fastest and ghost cells in place
    assert tuple(s._q.cur.shape) == (3, 8, 10) and s._q.pitch == 10 and s._q.mstride == 80
    assert float(s._q.cur[0, 2 + 3, 2 + 5]) == 23.0
    qbc = s.get_qbc_from_q(2, 'q')
    assert tuple(qbc.shape) == (3, 10, 8) and qbc.data_ptr() == s._q.cur.data_ptr()
    # numpy interop used by the reference's verifiers
    assert np.linalg.norm(np.asarray(s.q[0]) - np.arange(24.).reshape(6, 4)) == 0.0
    assert (np.arange(24.).reshape(6, 4) - s.q[0]).sum() == 0.0


def test_state_deepcopy_is_independent():
    s = _state2d(maux=2)
    s.q[...] = 1.0
    s.aux[...] = 2.0
    s.aux_global['g'] = 9.8
    c = copy.deepcopy(pyclaw.Solution(s))
    s.q[...] = 5.0
    assert float(c.q.sum()) == 3 * 6 * 4 and float(c.aux.sum()) == 2 * 2 * 6 * 4
    assert c.state.aux_global == {'g': 9.8} and c.state.grid is not s.grid


def test_ping_pong_backup_semantics():
    s = _state2d()
    s.set_mbc(2)
    s.q[...] = 1.0
    s._begin_step()
    new = s._q.get_spare()
    new.fill_(2.0)
    s._commit(new)
    assert float(s.q[0, 0, 0]) == 2.0
    s._reject_step()                       # rejected: roll back for free
    assert float(s.q[0, 0, 0]) == 1.0
    s._begin_step(copy=True)               # a hook may change q in place -> explicit copy
    s.q[...] = 7.0
    new = s._q.get_spare()
    new.fill_(3.0)
    s._commit(new)
    s._reject_step()
    assert float(s.q[0, 0, 0]) == 1.0
    s._begin_step()
    s._reject_step()                       # nothing committed (SharpClaw CFL error): no-op
    assert float(s.q[0, 0, 0]) == 1.0


def test_riemann_resolution():
    assert riemann.resolve(None, dict(rho=1, bulk=1, cc=1, zz=1), 2) is riemann.acoustics
    assert riemann.resolve(None, dict(gamma=1.4, gamma1=.4), 2) is riemann.euler_5wave
    assert riemann.resolve(None, dict(grav=1.), 2) is riemann.shallow_roe_with_efix
    assert riemann.resolve('advection', {}, 1) is riemann.advection
    assert riemann.resolve(riemann.rp_acoustics.rp_acoustics_2d, {}, 2) is riemann.acoustics
    with pytest.raises(NotImplementedError):
        riemann.resolve(lambda *a: None, {}, 1)      # Python Riemann solvers: no CPU path
    with pytest.raises(Exception, match="cparam"):
        riemann.euler_5wave.params(dict(gamma=1.4))


def test_solver_defaults_and_method_array():
    c = pyclaw.ClawSolver2D()
    assert (c.mbc, c.order, c.dim_split, c.order_trans, c.cfl_max, c.cfl_desired) == (2, 2, True, 1, 1.0, 0.9)
    assert c.limiters == pyclaw.limiters.tvd.minmod and c.src_split == 1 and c.kernel_language == 'Fortran'
    s = pyclaw.SharpClawSolver2D()
    assert (s.mbc, s.cfl_max, s.cfl_desired, s.time_integrator, s.lim_type, s.weno_order) == (3, 2.5, 2.45, 'SSP104', 2, 5)
    st = _state2d()
    c.mwaves = 2
    c.limiters = 4
    c.set_mthlim()
    assert c.mthlim == [4, 4]
    c.dim_split = False
    c.order_trans = 2
    c.set_method(st)
    assert c.method == [1, 2, 2, 0, 0, 0, 0]             # clawpack.py:192-212
    c.dim_split = True
    c.set_method(st)
    assert c.method[2] == -1
    c.limiters = [1, 2, 3]
    with pytest.raises(Exception):
        c.set_mthlim()
    assert pyclaw.BC.custom == 0 and pyclaw.BC.outflow == 1 and pyclaw.BC.periodic == 2 and pyclaw.BC.reflecting == 3


class _ScriptedSolver(Solver):
    """A Solver whose step() only reports a scripted Courant number: exercises the dt loop."""

    def __init__(self, cfls):
        self.ndim = 1
        self._required_attrs = list(Solver._base_required)
        d = dict(Solver._base_defaults)
        d.update(mbc=2, cfl_max=1.0, cfl_desired=0.9)
        self._default_attr_values = d
        super().__init__()
        self.cfls, self.log = list(cfls), []

    def _needs_backup_copy(self):
        return False

    def step(self, solution):
        # cfl proportional to dt, like a real solver: cfl = speed * dt
        self.log.append(self.dt)
        speed = self.cfls.pop(0)
        self.cfl.update_global_max(speed * self.dt)
        new = solution.state._q.get_spare()
        new.copy_(solution.state._q.cur)
        new += 1.0
        solution.state._commit(new)


def _sol():
    x = pyclaw.Dimension('x', 0., 1., 8)
    s = pyclaw.State(pyclaw.Grid(x), 1, device='cpu')
    s.set_mbc(2)
    return pyclaw.Solution(s)


def test_evolve_dt_schedule_and_final_step_rule():
    # solver.py:684-685: the step that reaches tend does NOT update dt
    sol = _sol()
    sv = _ScriptedSolver([10.0] * 20)
    sv.dt = 0.04
    st = sv.evolve_to_time(sol, 0.1)
    # dt: 0.04 (cfl .4) -> 0.09 -> clipped to 0.06 (reaches tend, no update)
    assert np.allclose(sv.log, [0.04, 0.06])
    assert st['numsteps'] == 2 and abs(sol.t - 0.1) < 1e-15
    assert sv.dt == pytest.approx(0.06)                   # carried into the next interval
    assert float(sol.q[0, 0]) == 2.0


def test_evolve_rejects_and_retries():
    sol = _sol()
    sv = _ScriptedSolver([10.0] * 20)
    sv.dt = 0.2                                           # cfl = 2 > cfl_max: rejected
    st = sv.evolve_to_time(sol, 0.3)
    assert sv.log[0] == 0.2 and sv.log[1] == pytest.approx(0.09)
    assert st['numsteps'] == len(sv.log) - 1
    assert float(sol.q[0, 0]) == st['numsteps']           # the rejected update was rolled back
    assert abs(sol.t - 0.3) < 1e-15


def test_evolve_fixed_dt():
    sol = _sol()
    sv = _ScriptedSolver([1.0] * 20)
    sv.dt_variable = False
    sv.dt = 0.05
    st = sv.evolve_to_time(sol, 0.2)
    assert st['numsteps'] == 4 and sol.t == pytest.approx(0.2)
    sv2 = _ScriptedSolver([100.0] * 4)
    sv2.dt_variable = False
    sv2.dt = 0.05
    with pytest.raises(Exception, match="CFL too large"):
        sv2.evolve_to_time(_sol(), 0.2)
    sv3 = _ScriptedSolver([1.0] * 4)
    sv3.dt_variable = False
    sv3.dt = 0.07
    with pytest.raises(Exception, match="does not divide"):
        sv3.evolve_to_time(_sol(), 0.2)


def test_evolve_max_steps_and_single_step():
    sv = _ScriptedSolver([1.0] * 50)
    sv.max_steps = 3
    sv.dt = 0.001
    sv.dt_max = 0.001
    with pytest.raises(Exception, match="Maximum number of timesteps"):
        sv.evolve_to_time(_sol(), 1.0)
    sol = _sol()
    sv = _ScriptedSolver([1.0] * 5)
    sv.dt = 0.01
    st = sv.evolve_to_time(sol)                           # tend=None: exactly one step
    assert st['numsteps'] == 1 and sol.t == pytest.approx(0.01)


def test_controller_output_times_and_frames():
    sol = _sol()
    sv = _ScriptedSolver([1.0] * 1000)
    sv.dt_initial = 0.03
    sv.bc_lower[0] = sv.bc_upper[0] = pyclaw.BC.periodic
    claw = pyclaw.Controller()
    claw.solution, claw.solver = sol, sv
    claw.tfinal, claw.nout, claw.keep_copy, claw.output_format = 0.3, 3, True, None
    status = claw.run()
    assert len(claw.frames) == 4
    assert [round(f.t, 12) for f in claw.frames] == [0.0, 0.1, 0.2, 0.3]
    assert status['numsteps'] >= 1
    # frames are snapshots, not views
    assert float(claw.frames[0].q[0, 0]) == 0.0 and float(claw.frames[-1].q[0, 0]) > 0.0


def test_data_container_round_trip(tmp_path):
    """pyclaw.Data (data.py:68-330): ``value =: name`` files."""
    import pyclaw
    f = tmp_path / "setprob.data"
    f.write_text("   1.4d0   =: gamma   # ratio of specific heats\n  3 =: mthlim_count\n 1 2 4 =: mthlim\n T =: efix\n shock =: name\n")
    d = pyclaw.Data(str(f))
    assert d.gamma == 1.4 and d.mthlim_count == 3 and d.mthlim == [1, 2, 4] and d.efix is True and d.name == 'shock'
    assert d.attributes == ['gamma', 'mthlim_count', 'mthlim', 'efix', 'name'] and d.get_owner('gamma') == str(f)
    d.add_attribute('cfl', 0.9)
    out = tmp_path / "out.data"
    d.write(str(out))
    e = pyclaw.Data(str(out))
    assert e.gamma == 1.4 and e.cfl == 0.9 and e.mthlim == [1, 2, 4] and e.efix is True
    d.remove_attributes('cfl')
    assert not d.has_attribute('cfl')
    with pytest.raises(NotImplementedError):
        pyclaw.plot.interactive_plot()
