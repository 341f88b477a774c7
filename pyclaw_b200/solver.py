"""Solver base class (src/pyclaw/solver.py:31-742): boundary conditions on the padded
device arrays and the dt / CFL accept-reject loop."""
import ctypes
import logging

import torch

from . import _lib
from .cfl import CFL


class CFLError(Exception):
    """Error raised when cfl_max is exceeded (solver.py:11-15)."""
    pass


class BC():
    """solver.py:17-23"""
    custom = 0
    outflow = 1
    periodic = 2
    reflecting = 3


def default_compute_gauge_values(q, aux):
    return q


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Solver(object):
    r"""
    Pyclaw solver superclass (solver.py:31-125).  Attributes: dt, cfl, status,
    dt_variable, max_steps, bc_lower/bc_upper, aux_bc_lower/aux_bc_upper,
    user_bc_lower/user_bc_upper, user_aux_bc_lower/user_aux_bc_upper.
    """
    _base_required = ['dt_initial', 'dt_max', 'cfl_max', 'cfl_desired', 'max_steps', 'dt_variable', 'mbc']
    _base_defaults = {'dt_initial': 0.1, 'dt_max': 1e99, 'max_steps': 1000, 'dt_variable': True}

    def __init__(self, data=None, claw_package=None):
        self.logger = logging.getLogger('evolve')
        if not hasattr(self, '_required_attrs'):
            self._required_attrs = list(self._base_required)
        if not hasattr(self, '_default_attr_values'):
            self._default_attr_values = dict(self._base_defaults)
        for (k, v) in self._default_attr_values.items():
            self.__dict__.setdefault(k, v)
        if data is not None:
            for attr in self._required_attrs:
                if hasattr(data, attr):
                    setattr(self, attr, getattr(data, attr))
        self.dt = self._default_attr_values['dt_initial']
        self.cfl = self._make_cfl(self._default_attr_values['cfl_desired'])
        self.status = {'cflmax': self.cfl.get_cached_max(), 'dtmin': self.dt, 'dtmax': self.dt, 'numsteps': 0}
        self.bc_lower = [None] * self.ndim
        self.bc_upper = [None] * self.ndim
        self.aux_bc_lower = [None] * self.ndim
        self.aux_bc_upper = [None] * self.ndim
        self.user_bc_lower = None
        self.user_bc_upper = None
        self.user_aux_bc_lower = None
        self.user_aux_bc_upper = None
        self.compute_gauge_values = default_compute_gauge_values
        self.rp = None
        self.qbc = None
        self.auxbc = None
        self._problem = None
        self._cfl_dev = None
        self._cfl_host = None
        self._halo = None  # set by the slab partition (pyclaw_b200.parallel)
        # 'strict': -fmad=false build, bit for bit against the reference's arithmetic (default);
        # 'fma': the same kernels with mul+add contraction (profiles/README.md has the error table)
        self.arithmetic = 'strict'

    def _make_cfl(self, v):
        return CFL(v)

    # ---- validation / setup stubs (solver.py:205-262) ----
    def is_valid(self):
        valid = True
        for key in self._required_attrs:
            if key not in self.__dict__:
                self.logger.info('%s is not present.' % key)
                valid = False
        if any([bc == BC.custom for bc in self.bc_lower]) and self.user_bc_lower is None:
            valid = False
        if any([bc == BC.custom for bc in self.bc_upper]) and self.user_bc_upper is None:
            valid = False
        return valid

    def setup(self, solution):
        pass

    def teardown(self):
        pass

    def __str__(self):
        output = "Solver Status:\n"
        for (k, v) in self.status.items():
            output = "\n".join((output, "%s = %s" % (k.rjust(25), v)))
        return output

    def allocate_rk_stages(self, solution):
        """solver.py:266-291"""
        nregisters = {'Euler': 1, 'SSP33': 2, 'SSP104': 3}[self.time_integrator]
        state = solution.states[0]
        State = type(state)
        self._rk_stages = []
        for i in range(nregisters - 1):
            s = State(state.grid, state.meqn, 0, device=state.device)
            s.aux_global = state.aux_global
            s.set_mbc(self.mbc)
            s.t = state.t
            if state.maux > 0:
                s._aux = state._aux
            self._rk_stages.append(s)

    # ---- device plumbing ----
    def _setup_device(self, state, method=None, mthlim=None, weno_variant=0):
        from . import riemann
        if state.device.type != 'cuda':
            raise _lib.ClawB200Error("pyclaw_b200 computes on CUDA devices only (no CPU fallback); "
                                     "state lives on %s" % state.device)
        grid = state.grid
        fwave = bool(getattr(self, 'fwave', False))
        self._rp = riemann.resolve(self.rp, state.aux_global, grid.ndim, fwave=fwave)
        self._variant = self.arithmetic
        if self._rp.lib is not None:
            # a user-supplied solver lives in its own variant of the library (riemann.from_header)
            if self.arithmetic not in ('strict', 'fma'):
                raise _lib.ClawB200Error("solver.arithmetic must be 'strict' or 'fma'")
            self._variant = self._rp._variant(self.arithmetic)
        _lib.set_variant(self._variant)
        _lib.load()
        if self._rp.fwave != fwave:
            # the reference links an f-wave solver into classic*fw.so and a wave solver into
            # classic*.so (clawpack.py:221-222); mixing them gives wrong second-order terms
            raise Exception("solver.fwave = %s but Riemann solver %s returns %s"
                            % (fwave, self._rp.name, "f-waves" if self._rp.fwave else "waves"))
        if state.maux < self._rp.maux:
            raise Exception("Riemann solver %s reads %d aux components; state.maux = %d"
                            % (self._rp.name, self._rp.maux, state.maux))
        if grid.ndim not in self._rp.ndims:
            raise Exception("Riemann solver %s has no %d-D version" % (self._rp.name, grid.ndim))
        if state.meqn != self._rp.meqn(grid.ndim):
            raise Exception("state.meqn = %d does not match Riemann solver %s" % (state.meqn, self._rp.name))
        if self.mwaves != self._rp.nwaves(grid.ndim):
            raise Exception("solver.mwaves = %s does not match Riemann solver %s (%d)"
                            % (self.mwaves, self._rp.name, self._rp.nwaves(grid.ndim)))
        ng, d = grid.ng, grid.d
        self._mz = ng[2] if grid.ndim > 2 else 1
        self._dz = d[2] if grid.ndim > 2 else 1.0
        self._problem = _lib.make_problem(
            grid.ndim, state.meqn, self.mwaves, self.mbc, ng[0], ng[1] if grid.ndim > 1 else 1,
            d[0], d[1] if grid.ndim > 1 else 1.0, self._rp.rp_id, self._rp.params(state.aux_global),
            method=method, mthlim=mthlim, maux=state.maux, pitch=state._q.pitch,
            mstride=state._q.mstride, weno_variant=weno_variant)
        self._halo = state._partition
        if self._halo is not None:
            self._halo.check_thickness(self.mbc)
        # a new problem invalidates every captured launch sequence (clawpack.py: the graphs hold
        # a byte copy of the previous problem and the previous buffers' addresses)
        self._dt_dev = None
        self._gproblem = None
        self._graphs = {}
        self._cfl_dev = torch.zeros(16, dtype=torch.float64, device=state.device)
        self._cfl_host = torch.zeros(16, dtype=torch.float64).pin_memory()

    def _read_cfl(self, n=1):
        """Device -> pinned host copy of the Courant number(s); the one host sync of a step."""
        if self._halo is not None:
            self._halo.allreduce_max(self._cfl_dev)
        self._cfl_host.copy_(self._cfl_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._cfl_host[:n].tolist()

    # ---- boundary conditions (solver.py:297-596) ----
    def allocate_bc_arrays(self, state):
        state.set_mbc(self.mbc)
        self.qbc = state._q.padded()
        if state.maux > 0:
            self.apply_aux_bcs(state)
        else:
            self.auxbc = None

    def _fill(self, state, arr_field, arr_view, bc_lower, bc_upper, user_lower, user_upper, is_q,
              exchange=True, dims=None):
        grid = state.grid
        P = self._problem
        if exchange and self._halo is not None:
            self._halo.exchange(arr_field, arr_field.ncomp, periodic=[b == BC.periodic for b in bc_lower],
                                problem=self._halo_problem(arr_field))
        for idim, dim in enumerate(grid.dimensions):
            if dims is not None and idim not in dims:
                continue
            for side, bcs, user in ((0, bc_lower, user_lower), (1, bc_upper, user_upper)):
                on_boundary = (dim.nstart == 0) if side == 0 else (dim.nend == dim.n)
                if not on_boundary:
                    continue
                bc = bcs[idim]
                if bc == BC.custom:
                    user(state, dim, state.t, arr_view, self.mbc)
                elif bc == BC.periodic and not (dim.nstart == 0 and dim.nend == dim.n):
                    pass  # wrap-around comes with the halo exchange (solver.py:366-367)
                elif bc in (BC.outflow, BC.periodic, BC.reflecting):
                    negate = idim + 1 if (is_q and bc == BC.reflecting) else -1
                    if grid.ndim == 3:
                        _lib.call("clawb200_bc_fill3", ctypes.byref(P), self._mz, _ptr(arr_field.cur),
                                  arr_field.ncomp, idim, side, bc, negate, _stream())
                    else:
                        _lib.call("clawb200_bc_fill", ctypes.byref(P), _ptr(arr_field.cur), arr_field.ncomp,
                                  idim, side, bc, negate, _stream())
                elif bc is None:
                    raise Exception("One or more of the boundary conditions has not been specified.")
                else:
                    raise NotImplementedError("Boundary condition %s not implemented" % bc)

    def _halo_problem(self, field):
        """The problem struct for the halo pack / unpack kernels, when ``field`` has q's padded
        shape (the kernels take pitch / mstride from it); None selects plain tensor copies."""
        P = self._problem
        if P is None or field.cur.dim() != 3 or field.pitch != P.pitch or field.mstride != P.mstride:
            return None
        return P

    def apply_q_bcs(self, state, exchange=True, dims=None):
        """Fill the ghost cells of the state's padded q (solver.py:315-381): dimension by
        dimension, lower then upper, so corner values come out as in the reference."""
        self.qbc = state._q.padded()
        self._fill(state, state._q, self.qbc, self.bc_lower, self.bc_upper,
                   self.user_bc_lower, self.user_bc_upper, True, exchange=exchange, dims=dims)

    def apply_aux_bcs(self, state):
        """solver.py:456-506; done once in setup (aux is time independent by default)."""
        self.auxbc = state._aux.padded()
        self._fill(state, state._aux, self.auxbc, self.aux_bc_lower, self.aux_bc_upper,
                   self.user_aux_bc_lower, self.user_aux_bc_upper, False)

    # ---- evolution (solver.py:602-717) ----
    def _needs_backup_copy(self):
        return True

    def evolve_to_time(self, solution, tend=None):
        _lib.set_variant(getattr(self, '_variant', None) or self.arithmetic)
        take_one_step = tend is None
        tstart = solution.t
        self.status['cflmax'] = self.cfl.get_cached_max()
        self.status['dtmin'] = self.dt
        self.status['dtmax'] = self.dt
        self.status['numsteps'] = 0
        max_steps = self.max_steps
        if not self.dt_variable:
            if take_one_step:
                max_steps = 1
            else:
                max_steps = int((tend - tstart + 1e-10) / self.dt)
                if abs(max_steps * self.dt - (tend - tstart)) > 1e-5 * (tend - tstart):
                    raise Exception('dt does not divide (tend-tstart) and dt is fixed!')
        if self.dt_variable == 1 and self.cfl_desired > self.cfl_max:
            raise Exception('Variable time-stepping and desired CFL > maximum CFL')
        if not take_one_step and tend <= tstart:
            self.logger.info("Already at or beyond end time: no evolution required.")
            max_steps = 0

        for n in range(max_steps):
            state = solution.state
            if not take_one_step and solution.t + self.dt > tend and tstart < tend:
                self.dt = tend - solution.t
            if self.dt_variable:
                # The solvers update q out of place, so the previous buffer doubles as the
                # reference's q_backup (solver.py:660); an explicit copy is only taken when
                # a user hook may modify q in place before the hyperbolic update.
                state._begin_step(copy=self._needs_backup_copy())
                told = solution.t

            self.step(solution)

            cfl = self.cfl.get_cached_max()
            if cfl <= self.cfl_max:
                self.status['cflmax'] = max(cfl, self.status['cflmax'])
                if self.dt_variable:
                    solution.t += self.dt
                else:
                    solution.t = tstart + (n + 1) * self.dt
                state._accept_step()
                self.logger.debug("Step %i  CFL = %f   dt = %f   t = %f" % (n, cfl, self.dt, solution.t))
                self.write_gauge_values(solution)
                self.status['numsteps'] += 1
                if take_one_step or solution.t >= tend:
                    break
            else:
                self.logger.debug("Rejecting time step, CFL number too large")
                if self.dt_variable:
                    state._reject_step()
                    solution.t = told
                else:
                    self.status['cflmax'] = max(cfl, self.status['cflmax'])
                    raise Exception('CFL too large, giving up!')

            if self.dt_variable:
                if cfl > 0.0:
                    self.dt = min(self.dt_max, self.dt * self.cfl_desired / cfl)
                    self.status['dtmin'] = min(self.dt, self.status['dtmin'])
                    self.status['dtmax'] = max(self.dt, self.status['dtmax'])
                else:
                    self.dt = self.dt_max

        if self.dt_variable and not take_one_step and solution.t < tend \
                and self.status['numsteps'] == max_steps:
            raise Exception("Maximum number of timesteps have been taken")
        return self.status

    def step(self, solution):
        raise NotImplementedError("No stepping routine has been defined!")

    # ---- gauges (solver.py:731-741) ----
    def write_gauge_values(self, solution):
        grid = solution.state.grid
        for i, gauge in enumerate(grid.gauges):
            idx = (slice(None),) + tuple(gauge)
            aux = solution.state.aux[idx] if solution.state.aux is not None else None
            q = solution.state.q[idx]
            p = self.compute_gauge_values(q, aux)
            t = solution.t
            grid.gauge_files[i].write(str(t) + ' ' + ' '.join(str(float(j)) for j in p) + '\n')
