"""Slab partition over the GPUs of one box: replaces PetClaw's PETSc DMDA
(src/petclaw/state.py:199-291) and its CFL allreduce (src/petclaw/cfl.py:29-31).

One process per GPU (torch.distributed, NCCL).  The grid is cut into contiguous slabs
along the LAST dimension (y in 2-D): in the q[m][j][i] layout a halo of ``mbc`` rows is
``mbc*pitch`` contiguous doubles per component, every rank owns the full x extent, so
corner ghost cells come with the rows and x-boundary conditions stay rank-local.

Per step (per Runge-Kutta stage for SharpClaw):
    pack mbc interior rows per side -> isend/irecv with the two neighbours -> unpack into
    the ghost rows -> physical boundary conditions on the ranks that own a boundary
    (solver.py:357,371 semantics through dim.nstart / dim.nend) -> sweeps
    -> all_reduce(MAX) of the 8-byte Courant number.
Results are bit-identical for any number of slabs: every cell update is a pure function
of its neighbourhood and max is associative (the reference's own standard: its 6-rank
run is compared with the serial golden file at 1e-14, test/test_examples.py:264-277).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slab_range(n, rank, size):
    """Contiguous block of a dimension of n cells owned by ``rank`` (balanced to +-1)."""
    return (n * rank) // size, (n * (rank + 1)) // size


class SlabPartition(object):
    def __init__(self, grid, group=None):
        self.rank, self.size = world()
        self.group = group
        self.dim_index = grid.ndim - 1
        dim = grid.dimensions[self.dim_index]
        nstart, nend = slab_range(dim.n, self.rank, self.size)
        if nend - nstart < 1:
            raise Exception("more ranks (%d) than cells (%d) along the partitioned dimension"
                            % (self.size, dim.n))
        dim._set_range(nstart, nend)
        self.nloc = nend - nstart
        self.lower_nbr = self.rank - 1
        self.upper_nbr = self.rank + 1
        self._bufs = {}

    def check_thickness(self, mbc):
        """The exchange sends the ``mbc`` interior rows next to each slab face: a slab thinner
        than the ghost width would send ghost rows.  Called once the solver knows ``mbc``."""
        if self.size > 1 and self.nloc < mbc:
            raise Exception("slab of %d cells is thinner than the ghost width mbc = %d"
                            % (self.nloc, mbc))

    def _buffers(self, field, nrows):
        # a "row" of the partitioned (last) dimension: a cell in 1-D, a padded row in 2-D, a
        # padded plane in 3-D
        tail = tuple(field.cur.shape[2:])
        key = (field.ncomp, nrows, tail, field.cur.device)
        if key not in self._bufs:
            shape = (field.ncomp, nrows) + tail
            mk = lambda: torch.empty(shape, dtype=torch.float64, device=field.cur.device)
            self._bufs[key] = dict(send_lo=mk(), send_hi=mk(), recv_lo=mk(), recv_hi=mk())
        return self._bufs[key]

    @staticmethod
    def _rows(field, r0, n):
        t = field.cur
        return t[:, r0:r0 + n] if t.dim() == 3 else t[:, r0:r0 + n]

    def _pack(self, field, row0, buf, problem):
        """mbc padded rows starting at array row ``row0`` -> contiguous buffer [m][row][i]:
        the halo kernel of the C ABI for 2-D fields, a strided copy otherwise."""
        if problem is not None and field.cur.dim() == 3 and field.cur.is_cuda:
            import ctypes
            from . import _lib
            from .solver import _ptr, _stream
            _lib.call("clawb200_halo_pack", ctypes.byref(problem), _ptr(field.cur), field.ncomp, row0,
                      field.mbc, _ptr(buf), _stream())
        else:
            buf.copy_(field.cur[:, row0:row0 + field.mbc])

    def _unpack(self, field, row0, buf, problem):
        if problem is not None and field.cur.dim() == 3 and field.cur.is_cuda:
            import ctypes
            from . import _lib
            from .solver import _ptr, _stream
            _lib.call("clawb200_halo_unpack", ctypes.byref(problem), _ptr(field.cur), field.ncomp, row0,
                      field.mbc, _ptr(buf), _stream())
        else:
            field.cur[:, row0:row0 + field.mbc].copy_(buf)

    def exchange(self, field, ncomp, periodic, problem=None):
        """Fill the ghost rows shared with neighbouring slabs.  ``periodic`` is the list of
        per-dimension flags: with a periodic partitioned dimension rank 0 and rank P-1
        are neighbours (the DMDA is created periodic, petclaw/state.py:205-208).
        ``problem`` (the solver's clawb200_problem) selects the library's pack / unpack kernels."""
        if self.size == 1:
            return
        mbc = field.mbc
        t = field.cur
        nloc = t.shape[1] - 2 * mbc
        if nloc < mbc:
            raise Exception("slab of %d cells is thinner than the ghost width mbc = %d" % (nloc, mbc))
        wrap = bool(periodic[self.dim_index])
        lo = self.lower_nbr if self.lower_nbr >= 0 else (self.size - 1 if wrap else None)
        hi = self.upper_nbr if self.upper_nbr < self.size else (0 if wrap else None)
        b = self._buffers(field, mbc)
        # Order matters when both neighbours are the same rank (2 ranks, periodic): NCCL
        # matches the messages of a pair by posting order, so "my top rows" must pair with
        # the peer's "lower ghost" receive: upward traffic first, downward traffic second.
        ops = []
        if hi is not None:
            self._pack(field, nloc, b['send_hi'], problem)
            ops.append(dist.P2POp(dist.isend, b['send_hi'], hi, self.group))
        if lo is not None:
            ops.append(dist.P2POp(dist.irecv, b['recv_lo'], lo, self.group))
            self._pack(field, mbc, b['send_lo'], problem)
            ops.append(dist.P2POp(dist.isend, b['send_lo'], lo, self.group))
        if hi is not None:
            ops.append(dist.P2POp(dist.irecv, b['recv_hi'], hi, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if lo is not None:
            self._unpack(field, 0, b['recv_lo'], problem)
        if hi is not None:
            self._unpack(field, nloc + mbc, b['recv_hi'], problem)

    def allreduce_max(self, cfl_dev):
        if self.size > 1:
            dist.all_reduce(cfl_dev, op=dist.ReduceOp.MAX, group=self.group)

    def gather_interior(self, state):
        """Assemble the global q on every rank (tests / small outputs only)."""
        q = state.q.contiguous()
        if self.size == 1:
            return q
        parts = [None] * self.size
        dist.all_gather_object(parts, q.cpu().numpy(), group=self.group)
        import numpy as np
        return np.concatenate(parts, axis=self.dim_index + 1)
