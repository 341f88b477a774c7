"""State (src/pyclaw/state.py:10-236) with q and aux resident in HBM.

Storage is one padded structure-of-arrays tensor per field, ``[m][j][i]`` with i
fastest and ``mbc`` ghost cells on every side -- the layout the sweep kernels stream.
``state.q`` / ``state.aux`` are zero-copy strided *views* of the interior in the
reference's index order ``q[m,i,j]``, so user code indexes exactly as before.  Because
the interior is a view of the padded array, the reference's get_qbc_from_q /
set_q_from_qbc copies (state.py:171-206) disappear.
"""
import copy

import numpy as np
import torch

from .array import as_claw, default_device
from .grid import Grid


class _Field(object):
    """A padded SoA device array plus spare buffers for ping-pong updates."""

    def __init__(self, ncomp, ng, mbc, device):
        self.ncomp, self.ng, self.mbc, self.device = ncomp, list(ng), mbc, device
        self.cur = self._alloc()
        self.spare = []

    def _alloc(self):
        shape = [self.ncomp] + [n + 2 * self.mbc for n in reversed(self.ng)]
        return torch.zeros(shape, dtype=torch.float64, device=self.device)

    def get_spare(self):
        return self.spare.pop() if self.spare else self._alloc()

    def put_spare(self, t):
        # buffers of another padded shape (a previous ghost-cell width) are dropped, never reused
        if t.shape == self.cur.shape and len(self.spare) < 3:
            self.spare.append(t)

    @staticmethod
    def user_view(t):
        """[m][j][i] / [m][k][j][i] storage -> [m,i,j] / [m,i,j,k] indexing"""
        if t.dim() == 2:
            return as_claw(t)
        return as_claw(t.permute(0, 2, 1) if t.dim() == 3 else t.permute(0, 3, 2, 1))

    def padded(self, t=None):
        return self.user_view(self.cur if t is None else t)

    def interior(self, t=None):
        v = self.padded(t)
        mbc = self.mbc
        if mbc == 0:
            return v
        idx = (slice(None),) + (slice(mbc, -mbc),) * (v.dim() - 1)
        return v[idx]

    def set_mbc(self, mbc):
        if mbc == self.mbc:
            return
        old = self.interior().clone()
        self.mbc = mbc
        self.cur = self._alloc()
        self.spare = []
        self.interior()[...] = old

    @property
    def pitch(self):
        return self.cur.shape[-1]

    @property
    def mstride(self):
        return self.cur.stride(0)


class State(object):
    r"""
    Contains the current state on a particular grid, including q, t, and aux
    (state.py:10-33).  ``State(grid, meqn, maux=0)``.
    """

    def __init__(self, grid, meqn, maux=0, device=None):
        if not isinstance(grid, Grid):
            raise Exception("A PyClaw State object must be initialized with a PyClaw Grid object.")
        self.grid = grid
        self.device = torch.device(device) if device is not None else default_device()
        self._partition = self._make_partition(grid)
        self.p = None
        self.F = None
        self.aux_global = {}
        self.t = 0.
        self.mcapa = -1
        self._q = _Field(meqn, grid.ng, 0, self.device)
        self._aux = _Field(maux, grid.ng, 0, self.device) if maux > 0 else None
        self._backup = None
        self._explicit = None

    def _make_partition(self, grid):
        """Single-GPU State: no partition.  petclaw.State overrides this."""
        return getattr(grid, '_partition', None)

    # ---- q / aux as interior views ----
    @property
    def q(self):
        return self._q.interior()

    @q.setter
    def q(self, value):
        self._assign(self._q, value)

    @property
    def aux(self):
        return None if self._aux is None else self._aux.interior()

    @aux.setter
    def aux(self, value):
        if value is None:
            self._aux = None
            return
        if self._aux is None:
            ncomp = value.shape[0]
            self._aux = _Field(ncomp, self.grid.ng, self._q.mbc, self.device)
        self._assign(self._aux, value)

    def _assign(self, field, value):
        dst = field.interior()
        if isinstance(value, torch.Tensor):
            if value.data_ptr() == dst.data_ptr() and value.stride() == dst.stride():
                return
            dst.copy_(value)
        else:
            dst[...] = np.asarray(value)

    meqn = property(lambda self: self._q.ncomp)
    maux = property(lambda self: 0 if self._aux is None else self._aux.ncomp)

    @property
    def mp(self):
        return 0 if self.p is None else self.p.shape[0]

    @mp.setter
    def mp(self, mp):
        if self.p is not None:
            raise Exception('Cannot change state.mp after p is initialized.')
        self.p = self.new_array(mp)

    @property
    def mF(self):
        return 0 if self.F is None else self.F.shape[0]

    @mF.setter
    def mF(self, mF):
        if self.F is not None:
            raise Exception('Cannot change state.mF after F is initialized.')
        self.F = self.new_array(mF)

    def __str__(self):
        output = "  t=%s meqn=%s\n  " % (self.t, self.meqn)
        output += "  q.shape=%s" % str(tuple(self.q.shape))
        if self.aux is not None:
            output += " aux.shape=%s" % str(tuple(self.aux.shape))
        return output

    def is_valid(self):
        return self._q is not None and self.meqn > 0

    def set_cparam(self, fortran_module):
        """state.py:142-162 (kept for API compatibility; the Riemann-solver constants
        travel by value in the C ABI's problem struct)."""
        for k, v in self.aux_global.items():
            setattr(fortran_module, k, v)

    def set_mbc(self, mbc):
        """Re-allocate the padded storage for ``mbc`` ghost cells (the PetClaw State
        re-creates its DMDA with the real stencil width here, petclaw/state.py:271-290)."""
        if mbc != self._q.mbc:
            # ping-pong buffers of the old padded shape must not come back as spares
            self._backup = None
            self._explicit = None
        self._q.set_mbc(mbc)
        if self._aux is not None:
            self._aux.set_mbc(mbc)

    def set_q_from_qbc(self, mbc, qbc):
        """state.py:171-186; a no-op when qbc is this state's own padded array."""
        idx = (slice(None),) + (slice(mbc, -mbc),) * self.grid.ndim
        self.q = qbc[idx]

    def get_qbc_from_q(self, mbc, whichvec, qbc=None):
        """state.py:188-206: returns the padded array (a view, no copy)."""
        field = self._q if whichvec == 'q' else self._aux
        if field.mbc != mbc:
            self.set_mbc(mbc)
        return field.padded()

    # ---- ping-pong support for the solvers (replaces q_backup, solver.py:659-661,690) ----
    def _release_backup(self):
        if self._backup is not None:
            self._q.put_spare(self._backup)
            self._backup = None

    def _begin_step(self, copy=False):
        self._release_backup()
        if self._explicit is not None:
            self._q.put_spare(self._explicit)
            self._explicit = None
        if copy:
            b = self._q.get_spare()
            b.copy_(self._q.cur)
            self._explicit = b

    def _commit(self, new_cur):
        """Install ``new_cur`` (a padded buffer from ``_q.get_spare()``) as the current q.
        The previous buffer is kept until the step is accepted, so rejecting is free."""
        self._release_backup()
        self._backup = self._q.cur
        self._q.cur = new_cur

    def _reject_step(self):
        if self._explicit is not None:
            self._q.put_spare(self._q.cur)
            self._q.cur = self._explicit
            self._explicit = None
            self._release_backup()
        elif self._backup is not None:
            self._q.put_spare(self._q.cur)
            self._q.cur = self._backup
            self._backup = None

    def _accept_step(self):
        if self._explicit is not None:
            self._q.put_spare(self._explicit)
            self._explicit = None
        self._release_backup()

    def __deepcopy__(self, memo={}):
        g = copy.deepcopy(self.grid)
        g._partition = self._partition
        for d_new, d_old in zip(g.dimensions, self.grid.dimensions):
            d_new._set_range(d_old.nstart, d_old.nend)
        result = self.__class__(g, self.meqn, self.maux, device=self.device)
        result.t = copy.deepcopy(self.t)
        result.set_mbc(0)
        result.q = self.q
        if self.aux is not None:
            result.aux = self.aux
        result.aux_global = copy.deepcopy(self.aux_global)
        result.mcapa = self.mcapa
        return result

    def sum_F(self, i):
        return float(torch.sum(torch.abs(self.F[i, ...])))

    def new_array(self, dof):
        if dof == 0:
            return None
        shape = [dof] + list(reversed(self.grid.ng))
        t = torch.zeros(shape, dtype=torch.float64, device=self.device)
        return _Field.user_view(t)
