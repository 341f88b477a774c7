"""Classic Clawpack solvers (src/pyclaw/clawpack.py:24-558) on the sm_100a sweep kernels.

``ClawSolver1D/2D.step_hyperbolic`` call libclawb200.so through ctypes where the
reference imports the f2py modules ``classic1`` / ``classic2`` (clawpack.py:317-323,
532-552).  q never leaves the GPU; the only host synchronisation per step is the
8-byte Courant number that evolve_to_time needs to accept or reject the step.
"""
import ctypes

from . import _lib, limiters
from .solver import Solver, _ptr, _stream


class ClawSolver(Solver):
    r"""
    Generic classic Clawpack solver (clawpack.py:24-87).  Attributes: limiters (mthlim),
    order, src_split, fwave, step_src, start_step, kernel_language, verbosity.
    """

    def __init__(self, data=None):
        self._required_attrs = list(Solver._base_required) + \
            ['limiters', 'order', 'src_split', 'fwave', 'step_src', 'start_step']
        d = dict(Solver._base_defaults)
        d.update({'mbc': 2, 'limiters': limiters.tvd.minmod, 'order': 2, 'src_split': 1,
                  'fwave': False, 'step_src': None, 'start_step': None,
                  'kernel_language': 'Fortran', 'verbosity': 0, 'cfl_max': 1.0, 'cfl_desired': 0.9})
        d.update(getattr(self, '_extra_defaults', {}))
        self._default_attr_values = d
        super(ClawSolver, self).__init__(data)

    # ---- clawpack.py:114-165 ----
    def step(self, solution):
        _lib.set_variant(getattr(self, '_variant', None) or self.arithmetic)
        if self.start_step is not None:
            self.start_step(self, solution)
        if self.src_split == 2 and self.step_src is not None:
            self.step_src(self, solution.states[0], self.dt / 2.0)
        self.step_hyperbolic(solution)
        # a step that will be rejected skips the source term (clawpack.py:153-154)
        if self.cfl.get_cached_max() >= self.cfl_max:
            return False
        if self.step_src is not None:
            if self.src_split == 2:
                self.step_src(self, solution.states[0], self.dt / 2.0)
            if self.src_split == 1:
                self.step_src(self, solution.states[0], self.dt)
        return True

    def _needs_backup_copy(self):
        return self.start_step is not None or (self.src_split == 2 and self.step_src is not None)

    # ---- CUDA-graph replay of the hyperbolic step -------------------------------------
    # A step is a fixed sequence of launches (ghost-cell fills, Courant-number reset, sweeps,
    # Courant-number read-back); only dt and the buffer addresses change.  dt is read by the
    # kernels from a device scalar (clawb200_problem.dt_dev), so the sequence is captured once
    # per buffer rotation and replayed: one graph launch per step instead of ~10 Python/ctypes
    # round trips.  Not used with custom (Python) boundary conditions or a slab partition.
    use_cuda_graph = True

    def _graph_ok(self, state):
        from .solver import BC
        return (self.use_cuda_graph and self._halo is None and
                BC.custom not in list(self.bc_lower) + list(self.bc_upper))

    def _hyperbolic_sequence(self, state, launch):
        """bc fill + cfl reset + sweeps + cfl read-back, eagerly or through a cached graph.
        ``launch(problem_ref, cfl_ptr, stream_ptr)`` issues the sweep kernels."""
        import torch
        if not self._graph_ok(state):
            self.apply_q_bcs(state)
            st = _stream()
            _lib.call("clawb200_cfl_reset", _ptr(self._cfl_dev), st)
            launch(ctypes.byref(self._problem), _ptr(self._cfl_dev), st)
            return self._read_cfl()[0]
        if getattr(self, '_dt_dev', None) is None:
            self._dt_dev = torch.zeros(1, dtype=torch.float64, device=state.device)
            self._dt_pin = torch.zeros(1, dtype=torch.float64).pin_memory()
            self._gproblem = type(self._problem).from_buffer_copy(self._problem)
            self._gproblem.dt_dev = self._dt_dev.data_ptr()
            self._graphs = {}

        def sequence():
            self._dt_dev.copy_(self._dt_pin, non_blocking=True)
            self.apply_q_bcs(state)
            st = _stream()
            _lib.call("clawb200_cfl_reset", _ptr(self._cfl_dev), st)
            launch(ctypes.byref(self._gproblem), _ptr(self._cfl_dev), st)
            self._cfl_host.copy_(self._cfl_dev, non_blocking=True)

        self._dt_pin[0] = float(self.dt)
        # everything a captured sequence depends on besides dt: the buffers, the aux array and
        # the boundary-condition types (the problem itself is covered by _setup_device)
        key = tuple(self._graph_key) + (
            state._aux.cur.data_ptr() if state._aux is not None else 0,
            tuple(self.bc_lower), tuple(self.bc_upper))
        g = self._graphs.get(key)
        if g is None:
            sequence()                               # this step, eagerly
            torch.cuda.current_stream().synchronize()
            cfl = float(self._cfl_host[0])
            if len(self._graphs) < 16:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):            # record the same sequence for next time
                    sequence()
                self._graphs[key] = g
            return cfl
        g.replay()
        torch.cuda.current_stream().synchronize()
        return float(self._cfl_host[0])

    def check_cfl_settings(self):
        pass

    def step_hyperbolic(self, solution):
        raise Exception("Dummy routine, please override!")

    def set_mthlim(self):
        """clawpack.py:181-190"""
        self.mthlim = self.limiters
        if not isinstance(self.limiters, list):
            self.mthlim = [self.mthlim]
        if len(self.mthlim) == 1:
            self.mthlim = self.mthlim * self.mwaves
        if len(self.mthlim) != self.mwaves:
            raise Exception('Length of solver.limiters is not equal to 1 or to solver.mwaves')

    def set_method(self, state):
        """The Fortran ``method`` array (clawpack.py:192-212)."""
        self.method = [0] * 7
        self.method[0] = int(self.dt_variable)
        self.method[1] = self.order
        if self.ndim == 1:
            self.method[2] = 0
        elif self.dim_split:
            self.method[2] = -1
        else:
            self.method[2] = self.order_trans
        self.method[3] = self.verbosity
        self.method[4] = 0
        self.method[5] = state.mcapa + 1
        self.method[6] = state.maux

    def setup(self, solution):
        """clawpack.py:214-238"""
        if self.kernel_language not in ('Fortran', 'CUDA'):
            raise NotImplementedError("kernel_language=%r: only the CUDA kernels exist ('Fortran' is "
                                      "accepted as an alias); there is no Python/CPU path" % self.kernel_language)
        state = solution.state
        state.set_mbc(self.mbc)
        self.check_cfl_settings()
        self.set_mthlim()
        self.set_method(state)
        self._setup_device(state, method=self.method, mthlim=self.mthlim)
        self.allocate_bc_arrays(state)

    def teardown(self):
        pass


class ClawSolver1D(ClawSolver):
    """clawpack.py:271-406"""

    def __init__(self, data=None):
        self.ndim = 1
        super(ClawSolver1D, self).__init__(data)

    def step_hyperbolic(self, solution):
        state = solution.states[0]
        qold = state._q.cur
        qnew = state._q.get_spare()
        aux = _ptr(state._aux.cur if state._aux is not None else None)
        dt = float(self.dt)

        def launch(P, cfl, st):
            _lib.call("clawb200_step1", P, _ptr(qold), _ptr(qnew), aux, dt, cfl, st)
        self._graph_key = (qold.data_ptr(), qnew.data_ptr())
        cfl = self._hyperbolic_sequence(state, launch)
        state._commit(qnew)
        self.cfl.update_global_max(cfl)


class ClawSolver2D(ClawSolver):
    """clawpack.py:411-558.  dim_split / order_trans as in the reference."""
    no_trans = 0
    trans_inc = 1
    trans_cor = 2

    def __init__(self, data=None):
        self._extra_defaults = {'dim_split': True, 'order_trans': self.trans_inc}
        self.ndim = 2
        super(ClawSolver2D, self).__init__(data)

    def check_cfl_settings(self):
        if (not self.dim_split) and (self.order_trans == 0):
            cfl_recommended = 0.5
        else:
            cfl_recommended = 1.0
        if self.cfl_max > cfl_recommended:
            import warnings
            warnings.warn('cfl_max is set higher than the recommended value of %s' % cfl_recommended)

    # Multi-GPU (slab partition), unsplit: the halo exchange runs on its own stream while
    # the rows that do not depend on a neighbour's halo are updated; the mbc boundary rows
    # per side follow once the halo has landed.
    overlap_halo = True

    def _step_overlapped(self, state):
        """Two streams.  Side stream: halo exchange -> ghost rows' boundary conditions -> the
        ``mbc`` boundary rows per side (they need the halo).  Main stream: x-direction boundary
        conditions -> the interior rows (they do not).  The boundary bands are small launches
        (2 * mbc rows: less than one wave of CTAs); running them beside the interior sweeps
        instead of after them takes their latency off the step."""
        import torch
        from .solver import BC
        part, F, mbc = self._halo, state._q, self.mbc
        my = state.grid.ng[1]
        qold, qnew = F.cur, F.get_spare()
        aux = _ptr(state._aux.cur if state._aux is not None else None)
        dt = float(self.dt)
        P, cfl = ctypes.byref(self._problem), _ptr(self._cfl_dev)
        cur = torch.cuda.current_stream()
        if getattr(self, '_hstream', None) is None:
            # high priority: the small boundary launches must not queue behind the interior sweeps' CTAs
            self._hstream = torch.cuda.Stream(priority=-1)
            self._ev0, self._ev1, self._evx = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        _lib.call("clawb200_cfl_reset", cfl, _stream())
        self._ev0.record(cur)
        # main stream: x-direction boundary conditions of the rows this rank already has, then the
        # interior sweeps -- enqueued BEFORE the side stream's ~30 small operations, so that the
        # GPU is not idle while the host is busy issuing them
        self.apply_q_bcs(state, exchange=False, dims=[0])
        self._evx.record(cur)
        _lib.call("clawb200_step2_rows", P, _ptr(qold), _ptr(qnew), aux, dt, 1 + mbc, my - mbc, cfl, _stream())
        with torch.cuda.stream(self._hstream):
            self._hstream.wait_event(self._ev0)
            part.exchange(F, F.ncomp, periodic=[b == BC.periodic for b in self.bc_lower],
                          problem=self._halo_problem(F))
            # the halo has landed: ghost rows get their x-BCs (after the main stream's fill, which
            # saw them half-written), edge ranks their y-BCs, then the boundary rows are updated
            self._hstream.wait_event(self._evx)
            self.apply_q_bcs(state, exchange=False)
            st = _stream()
            _lib.call("clawb200_step2_rows", P, _ptr(qold), _ptr(qnew), aux, dt, 1, mbc, cfl, st)
            _lib.call("clawb200_step2_rows", P, _ptr(qold), _ptr(qnew), aux, dt, my - mbc + 1, my, cfl, st)
            self._ev1.record(self._hstream)
        cur.wait_event(self._ev1)
        state._commit(qnew)
        self.cfl.update_global_max(self._read_cfl()[0])

    def step_hyperbolic(self, solution):
        state = solution.states[0]
        if (self._halo is not None and self._halo.size > 1 and self.overlap_halo and not self.dim_split
                and state.grid.ng[1] > 4 * self.mbc):
            return self._step_overlapped(state)
        aux = _ptr(state._aux.cur if state._aux is not None else None)
        dt = float(self.dt)
        qold = state._q.cur
        qnew = state._q.get_spare()
        tmp = state._q.get_spare() if self.dim_split else None

        def launch(P, cfl, st):
            if self.dim_split:
                # step2ds twice (clawpack.py:538-548); the Fortran's in-place second call is a
                # ping-pong here: qold -> tmp (x-sweeps) -> qnew (y-sweeps)
                _lib.call("clawb200_step2ds", P, _ptr(qold), _ptr(tmp), aux, dt, 1, cfl, st)
                _lib.call("clawb200_step2ds", P, _ptr(tmp), _ptr(qnew), aux, dt, 2, cfl, st)
            else:
                _lib.call("clawb200_step2", P, _ptr(qold), _ptr(qnew), aux, dt, cfl, st)
        self._graph_key = (qold.data_ptr(), qnew.data_ptr(), tmp.data_ptr() if tmp is not None else 0)
        cfl = self._hyperbolic_sequence(state, launch)
        if tmp is not None:
            state._q.put_spare(tmp)
        state._commit(qnew)
        self.cfl.update_global_max(cfl)


class ClawSolver3D(ClawSolver):
    """clawpack.py:563-702: the dimensionally split algorithm (``dim_split=True``, the reference's
    default: three ``step3ds`` sweeps) and the unsplit one (``step3`` with the transverse and
    double-transverse solves selected by ``order_trans``)."""
    no_trans = 0
    trans_inc = 11
    trans_cor = 22

    def __init__(self, data=None):
        self._extra_defaults = {'dim_split': True, 'order_trans': self.trans_cor}
        self.ndim = 3
        super(ClawSolver3D, self).__init__(data)

    def step_hyperbolic(self, solution):
        state = solution.states[0]
        aux = _ptr(state._aux.cur if state._aux is not None else None)
        dt = float(self.dt)
        F = state._q
        qold, b1, b2 = F.cur, F.get_spare(), F.get_spare()
        mz, dz = int(self._mz), float(self._dz)
        if not self.dim_split:
            import torch
            need = _lib.load().clawb200_step3_scratch_doubles(ctypes.byref(self._problem))
            if getattr(self, '_scratch3', None) is None or self._scratch3.numel() < need:
                self._scratch3 = torch.empty(need, dtype=torch.float64, device=state.device)
            scratch = _ptr(self._scratch3)

        def launch(P, cfl, st):
            if not self.dim_split:
                # classic3.step3 (clawpack.py:680-682)
                _lib.call("clawb200_step3", P, mz, dz, _ptr(qold), _ptr(b1), aux, dt, scratch, cfl, st)
                return
            # step3ds three times (clawpack.py:656-676); the Fortran's aliased calls become a
            # rotation qold -> b1 (x) -> b2 (y) -> b1 (z)
            _lib.call("clawb200_step3ds", P, mz, dz, _ptr(qold), _ptr(b1), aux, dt, 1, cfl, st)
            _lib.call("clawb200_step3ds", P, mz, dz, _ptr(b1), _ptr(b2), aux, dt, 2, cfl, st)
            _lib.call("clawb200_step3ds", P, mz, dz, _ptr(b2), _ptr(b1), aux, dt, 3, cfl, st)
        self._graph_key = (qold.data_ptr(), b1.data_ptr(), b2.data_ptr())
        cfl = self._hyperbolic_sequence(state, launch)
        F.put_spare(b2)
        state._commit(b1)
        self.cfl.update_global_max(cfl)
